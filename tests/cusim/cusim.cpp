// tests/cusim -- the fibre machine and the runtime-API subset behind tests/cusim/include/*.h.
// TEST INFRASTRUCTURE (see the header of include/cuda_runtime.h): a CPU emulation of the CUDA execution model that runs
// hypercraft_b200/csrc/*.cu unchanged.  Never part of the product, never measured.
//
// Scheduling.  A launch is executed synchronously: worker OS threads pull CTA indices from a counter; a CTA's threads are
// fibres on that worker.  Within a CTA the scheduler runs ONE WARP AT A TIME for as long as any of its lanes can move
// (lanes stop at warp collectives, at __syncthreads, at the end of the kernel, or in a spin-wait), then the next warp; when
// every live fibre is parked at __syncthreads the barrier opens.  That is the most skewed legal schedule there is -- a
// missing barrier between a producer and a consumer warp shows up as stale data instead of being hidden by lock-step timing
// -- and CUSIM_REVERSE=1 visits warps and lanes in descending order to expose the opposite direction.
//
// Environment: CUSIM_SMS (multiProcessorCount, default 8), CUSIM_THREADS (worker OS threads, default min(cores, 16)),
// CUSIM_DEVICES (device count, default 1), CUSIM_REVERSE, CUSIM_EVENT_BUSY (n: every cudaEventQuery answers "not ready" n
// times before it answers "done" -- lets host code that polls the copy engine see a busy engine).
#include <cuda.h>
#include <cuda_runtime.h>

#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

#include <sys/mman.h>

namespace cusim {
namespace {

[[noreturn]] void die(const char *fmt, ...) __attribute__((format(printf, 1, 2)));
void die(const char *fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	fprintf(stderr, "[cusim] fatal: ");
	vfprintf(stderr, fmt, ap);
	fprintf(stderr, "\n");
	va_end(ap);
	abort();
}

int env_int(const char *name, int dflt) {
	const char *v = getenv(name);
	return v && *v ? atoi(v) : dflt;
}

// ---- context switch (x86-64 System V: callee-saved registers + stack pointer) -------------------------------------------
extern "C" void cusim_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.globl cusim_switch
.type cusim_switch,@function
cusim_switch:
	pushq %rbp
	pushq %rbx
	pushq %r12
	pushq %r13
	pushq %r14
	pushq %r15
	movq %rsp, (%rdi)
	movq %rsi, %rsp
	popq %r15
	popq %r14
	popq %r13
	popq %r12
	popq %rbx
	popq %rbp
	ret
.size cusim_switch,.-cusim_switch
)");

constexpr size_t kStackBytes = 96 * 1024;
constexpr size_t kMaxDynSmem = 232448;

enum State : uint8_t { READY, WAIT_WARP, WAIT_CTA, SPIN, DONE };

struct WarpRec { // one rendez-vous group (keyed by its member mask) of one warp
	unsigned long long slot[2][32];
	unsigned arrived = 0, gen = 0;
};

struct Fiber {
	void *sp = nullptr;
	State state = DONE;
	ThreadState ts{};
	WarpRec *wait_rec = nullptr;
	unsigned wait_gen = 0;
	unsigned spin_pass = 0;
};

struct Worker { // one OS thread = one CTA at a time
	std::vector<Fiber> fibers;
	uint8_t *stacks = nullptr;
	uint8_t *smem = nullptr;
	void *sched_sp = nullptr;
	Fiber *cur = nullptr;
	const std::function<void()> *body = nullptr;
	std::vector<std::unordered_map<unsigned, WarpRec>> recs; // per warp
	unsigned n_threads = 0, n_alive = 0, n_at_barrier = 0;
	int bar_or = 0, bar_and = 1, bar_result_or = 0, bar_result_and = 1;
	unsigned pass = 0;
	~Worker() {
		if (stacks)
			munmap(stacks, kStackBytes * 1024);
		free(smem);
	}
};
thread_local Worker *tl_worker = nullptr;
thread_local ThreadState tl_host_state{}; // threadIdx etc. read outside a kernel (never, but keep it defined)

void to_scheduler() {
	Worker *w = tl_worker;
	cusim_switch(&w->cur->sp, w->sched_sp);
}

extern "C" void cusim_fiber_main() {
	Worker *w = tl_worker;
	(*w->body)();
	w->cur->state = DONE;
	--w->n_alive;
	to_scheduler();
	die("a finished fibre was resumed");
}

void prepare_fiber(Worker *w, unsigned t) {
	uint8_t *top = w->stacks + (size_t)(t + 1) * kStackBytes;
	void **sp = reinterpret_cast<void **>(top);
	*--sp = nullptr;                                  // fake return address of cusim_fiber_main (keeps rsp = 8 mod 16 at its entry)
	*--sp = reinterpret_cast<void *>(cusim_fiber_main); // popped by `ret`
	for (int i = 0; i < 6; ++i)
		*--sp = nullptr; // r15 r14 r13 r12 rbx rbp
	w->fibers[t].sp = sp;
}

void run_block(Worker *w, dim3 grid, dim3 block, unsigned bx, const std::function<void()> &body, bool reverse) {
	const unsigned n = block.x * block.y * block.z;
	if (n == 0 || n > 1024)
		die("block of %u threads", n);
	if (!w->stacks) {
		w->stacks = (uint8_t *)mmap(nullptr, kStackBytes * 1024, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
		if (w->stacks == MAP_FAILED)
			die("cannot map fibre stacks");
		if (posix_memalign((void **)&w->smem, 1024, kMaxDynSmem))
			die("cannot allocate shared memory");
	}
	memset(w->smem, 0xcd, kMaxDynSmem); // shared memory of a fresh CTA holds garbage
	w->fibers.assign(n, Fiber{});
	const unsigned n_warps = (n + 31) / 32;
	w->recs.assign(n_warps, {});
	w->n_threads = n, w->n_alive = n, w->n_at_barrier = 0, w->bar_or = 0, w->bar_and = 1, w->body = &body, w->pass = 0;
	for (unsigned t = 0; t < n; ++t) {
		Fiber &f = w->fibers[t];
		f.state = READY;
		f.ts.tid = Idx{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
		f.ts.bid = Idx{bx, 0, 0};
		f.ts.bdim = Idx{block.x, block.y, block.z};
		f.ts.gdim = Idx{grid.x, grid.y, grid.z};
		f.ts.lane = t & 31, f.ts.warp = t >> 5;
		prepare_fiber(w, t);
	}
	unsigned idle_passes = 0;
	while (w->n_alive) {
		++w->pass;
		bool progressed = false;
		for (unsigned wi = 0; wi < n_warps; ++wi) {
			const unsigned wp = reverse ? n_warps - 1 - wi : wi;
			for (;;) {
				bool any = false;
				for (unsigned li = 0; li < 32; ++li) {
					const unsigned t = wp * 32 + (reverse ? 31 - li : li);
					if (t >= n)
						continue;
					Fiber &f = w->fibers[t];
					switch (f.state) {
					case DONE:
					case WAIT_CTA:
						continue;
					case WAIT_WARP:
						if (f.wait_rec->gen == f.wait_gen)
							continue;
						break;
					case SPIN:
						if (f.spin_pass == w->pass)
							continue; // one poll per pass
						break;
					case READY:
						break;
					}
					const bool was_spin = f.state == SPIN;
					f.state = READY;
					w->cur = &f;
					cusim_switch(&w->sched_sp, f.sp);
					any = true;
					if (!(was_spin && f.state == SPIN))
						progressed = true;
				}
				if (!any)
					break;
			}
		}
		if (w->n_alive && w->n_at_barrier == w->n_alive) { // __syncthreads opens (exited threads do not take part)
			w->bar_result_or = w->bar_or, w->bar_result_and = w->bar_and;
			w->bar_or = 0, w->bar_and = 1, w->n_at_barrier = 0;
			for (Fiber &f : w->fibers)
				if (f.state == WAIT_CTA)
					f.state = READY;
			progressed = true;
		}
		if (progressed)
			idle_passes = 0;
		else if (++idle_passes > 100000) {
			unsigned c[5] = {0, 0, 0, 0, 0};
			for (Fiber &f : w->fibers)
				++c[f.state];
			die("deadlock in block %u: %u ready, %u at a warp collective, %u at __syncthreads, %u spinning, %u done (a collective names a lane that "
			    "exited or diverged, a barrier is not reached by every live thread, or an mbarrier never completes)",
			    bx, c[READY], c[WAIT_WARP], c[WAIT_CTA], c[SPIN], c[DONE]);
		}
	}
	w->cur = nullptr;
}

Fiber &cur_fiber() {
	Worker *w = tl_worker;
	if (!w || !w->cur)
		die("device intrinsic called outside a kernel");
	return *w->cur;
}

std::atomic<int> g_last_error{0};

} // namespace

ThreadState &self() {
	Worker *w = tl_worker;
	return w && w->cur ? w->cur->ts : tl_host_state;
}
unsigned char *dynamic_smem() { return tl_worker->smem; }

void launch(dim3 grid, dim3 block, size_t smem, std::function<void()> body) {
	if (smem > kMaxDynSmem)
		die("launch with %zu bytes of dynamic shared memory", smem);
	const unsigned n_blocks = grid.x * grid.y * grid.z;
	if (grid.y != 1 || grid.z != 1)
		die("only 1-D grids");
	if (tl_worker && tl_worker->cur)
		die("nested launch");
	const bool reverse = env_int("CUSIM_REVERSE", 0) != 0;
	const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
	const unsigned n_workers = std::min<unsigned>(n_blocks, (unsigned)std::max(1, env_int("CUSIM_THREADS", (int)std::min(hw, 16u))));
	std::atomic<unsigned> next{0};
	auto work = [&]() {
		Worker w;
		tl_worker = &w;
		for (;;) {
			const unsigned b = next.fetch_add(1);
			if (b >= n_blocks)
				break;
			run_block(&w, grid, block, b, body, reverse);
		}
		tl_worker = nullptr;
	};
	if (n_workers <= 1)
		work();
	else {
		std::vector<std::thread> ts;
		for (unsigned i = 0; i < n_workers; ++i)
			ts.emplace_back(work);
		for (auto &t : ts)
			t.join();
	}
}

void sync_threads() { (void)sync_threads_or(0); }
int sync_threads_or(int p) {
	Worker *w = tl_worker;
	Fiber &f = cur_fiber();
	w->bar_or |= p ? 1 : 0, w->bar_and &= p ? 1 : 0;
	++w->n_at_barrier;
	f.state = WAIT_CTA;
	to_scheduler();
	return w->bar_result_or;
}
int sync_threads_and(int p) {
	Worker *w = tl_worker;
	Fiber &f = cur_fiber();
	w->bar_or |= p ? 1 : 0, w->bar_and &= p ? 1 : 0;
	++w->n_at_barrier;
	f.state = WAIT_CTA;
	to_scheduler();
	return w->bar_result_and;
}
void yield() {
	Worker *w = tl_worker;
	Fiber &f = cur_fiber();
	f.state = SPIN, f.spin_pass = w->pass;
	to_scheduler();
}

const unsigned long long *warp_gather(unsigned mask, unsigned long long v) {
	Worker *w = tl_worker;
	Fiber &f = cur_fiber();
	const unsigned lane = f.ts.lane;
	if (!((mask >> lane) & 1))
		die("lane %u calls a warp collective with mask %08x that does not name it", lane, mask);
	const unsigned warp_threads = std::min(32u, w->n_threads - f.ts.warp * 32);
	if (warp_threads < 32 && (mask >> warp_threads))
		mask &= (1u << warp_threads) - 1u; // lanes beyond the block do not exist
	WarpRec &r = w->recs[f.ts.warp][mask];
	const unsigned g = r.gen;
	if ((r.arrived >> lane) & 1)
		die("lane %u arrives twice at a collective (mask %08x)", lane, mask);
	r.slot[g & 1][lane] = v;
	r.arrived |= 1u << lane;
	if (r.arrived == mask) {
		r.arrived = 0, ++r.gen;
	} else {
		f.state = WAIT_WARP, f.wait_rec = &r, f.wait_gen = g;
		to_scheduler();
	}
	return r.slot[g & 1];
}

// ---- mbarrier: bit 63 phase, bits 40..59 expected arrivals, bits 20..39 pending arrivals, bits 0..19 pending bytes + 2^19 -----
namespace {
constexpr unsigned long long kTxBias = 1ull << 19;
struct Mbar {
	unsigned phase, expected, pending;
	long long tx;
};
Mbar unpack(unsigned long long v) { return Mbar{(unsigned)(v >> 63), (unsigned)((v >> 40) & 0xfffff), (unsigned)((v >> 20) & 0xfffff), (long long)(v & 0xfffff) - (long long)kTxBias}; }
unsigned long long pack(const Mbar &m) {
	return ((unsigned long long)m.phase << 63) | ((unsigned long long)m.expected << 40) | ((unsigned long long)m.pending << 20) | (unsigned long long)(m.tx + (long long)kTxBias);
}
void settle(unsigned long long *bar, Mbar m) {
	if (m.tx <= -(long long)kTxBias || m.tx >= (long long)kTxBias)
		die("mbarrier transaction count out of range");
	if (m.pending == 0 && m.tx == 0)
		m.phase ^= 1, m.pending = m.expected;
	*bar = pack(m);
}
struct TensorMap {
	uint64_t magic;
	uint8_t *base;
	uint32_t elem, rank;
	uint64_t dims[3], strides[3]; // strides in bytes, strides[0] = elem
	uint32_t box[3];
};
static_assert(sizeof(TensorMap) <= sizeof(CUtensorMap), "tensor map fits");
constexpr uint64_t kMapMagic = 0x43555349'4d544d31ull;
} // namespace

void mbar_init(unsigned long long *bar, unsigned count) {
	if ((size_t)bar & 7)
		die("mbarrier not 8-byte aligned");
	*bar = pack(Mbar{0, count, count, 0});
}
void mbar_expect_tx(unsigned long long *bar, unsigned bytes) { // mbarrier.arrive.expect_tx
	Mbar m = unpack(*bar);
	if (m.pending == 0)
		die("mbarrier arrival beyond the expected count");
	m.tx += bytes, --m.pending;
	settle(bar, m);
}
static void complete_tx(unsigned long long *bar, unsigned bytes) {
	Mbar m = unpack(*bar);
	m.tx -= bytes;
	settle(bar, m);
}
void mbar_wait(unsigned long long *bar, unsigned parity) {
	while (unpack(*bar).phase == (parity & 1))
		yield();
}
void tma_load_1d(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
	if (((size_t)dst & 15) || ((size_t)src & 15) || (bytes & 15))
		die("cp.async.bulk: addresses and size must be multiples of 16 (dst %p src %p bytes %u)", dst, src, bytes);
	memcpy(dst, src, bytes);
	complete_tx(bar, bytes);
}
void tma_load_3d(void *dst, const void *map, int c0, int c1, int c2, unsigned long long *bar) {
	const TensorMap &m = *reinterpret_cast<const TensorMap *>(map);
	if (m.magic != kMapMagic || m.rank != 3)
		die("cp.async.bulk.tensor.3d with a tensor map this emulation did not encode");
	if ((size_t)dst & 127)
		die("cp.async.bulk.tensor: shared-memory destination %p not 128-byte aligned", dst);
	if (((long long)c0 * m.elem) & 15)
		die("cp.async.bulk.tensor: innermost start coordinate %d x %u bytes is not a multiple of 16 (measured on B200: illegal instruction)", c0, m.elem);
	uint8_t *d = (uint8_t *)dst;
	const int c[3] = {c0, c1, c2};
	const size_t row = (size_t)m.box[0] * m.elem;
	for (uint32_t k = 0; k < m.box[2]; ++k)
		for (uint32_t j = 0; j < m.box[1]; ++j, d += row) {
			const long long y = (long long)c[1] + j, z = (long long)c[2] + k;
			if (y < 0 || z < 0 || y >= (long long)m.dims[1] || z >= (long long)m.dims[2]) {
				memset(d, 0, row); // out-of-bounds elements read as zero
				continue;
			}
			const uint8_t *s = m.base + (size_t)z * m.strides[2] + (size_t)y * m.strides[1];
			for (uint32_t i = 0; i < m.box[0]; ++i) {
				const long long x = (long long)c[0] + i;
				if (x < 0 || x >= (long long)m.dims[0])
					memset(d + (size_t)i * m.elem, 0, m.elem);
				else
					memcpy(d + (size_t)i * m.elem, s + (size_t)x * m.elem, m.elem);
			}
		}
	complete_tx(bar, (unsigned)(row * m.box[1] * m.box[2]));
}

} // namespace cusim

// =================================================================================================================
// runtime API subset: everything is synchronous, "device memory" is host memory
// =================================================================================================================
using cusim::g_last_error;
struct cusimStream { int dummy; };
struct cusimEvent {
	std::chrono::steady_clock::time_point t;
	bool recorded = false;
	int busy_polls = 0;
};

extern "C" {

const char *cudaGetErrorString(cudaError_t e) {
	switch (e) {
	case cudaSuccess: return "no error";
	case cudaErrorInvalidValue: return "invalid argument";
	case cudaErrorMemoryAllocation: return "out of memory";
	case cudaErrorNotReady: return "device not ready";
	case cudaErrorNoDevice: return "no CUDA-capable device is detected";
	default: return "unknown error";
	}
}
cudaError_t cudaGetLastError() { return g_last_error.exchange(0); }
cudaError_t cudaGetDeviceCount(int *n) {
	*n = cusim::env_int("CUSIM_DEVICES", 1);
	return cudaSuccess;
}
cudaError_t cudaSetDevice(int d) { return d >= 0 && d < cusim::env_int("CUSIM_DEVICES", 1) ? cudaSuccess : cudaErrorInvalidValue; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
	memset(p, 0, sizeof(*p));
	snprintf(p->name, sizeof(p->name), "cusim (CPU emulation, test infrastructure)");
	p->major = 10, p->minor = 0, p->multiProcessorCount = std::max(1, cusim::env_int("CUSIM_SMS", 8));
	p->totalGlobalMem = (size_t)64 << 30, p->sharedMemPerBlockOptin = cusim::kMaxDynSmem;
	return cudaSuccess;
}
cudaError_t cudaMalloc(void **p, size_t n) {
	// device allocations are 256-byte aligned; fill with garbage so that reads of never-written device memory show up
	if (posix_memalign(p, 256, std::max<size_t>(n, 1)))
		return cudaErrorMemoryAllocation;
	if (n <= ((size_t)1 << 28))
		memset(*p, 0xa5, n);
	return cudaSuccess;
}
cudaError_t cudaFree(void *p) {
	free(p);
	return cudaSuccess;
}
cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { return posix_memalign(p, 4096, std::max<size_t>(n, 1)) ? cudaErrorMemoryAllocation : cudaSuccess; }
cudaError_t cudaFreeHost(void *p) {
	free(p);
	return cudaSuccess;
}
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) {
	memmove(d, s, n);
	return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) {
	memmove(d, s, n);
	return cudaSuccess;
}
cudaError_t cudaMemcpyBatchAsync(void **dsts, void **srcs, size_t *sizes, size_t count, cudaMemcpyAttributes *, size_t *, size_t, size_t *, cudaStream_t) {
	if (cusim::env_int("CUSIM_NO_BATCH_COPY", 0))
		return cudaErrorInvalidValue;
	for (size_t i = 0; i < count; ++i)
		memmove(dsts[i], srcs[i], sizes[i]);
	return cudaSuccess;
}
cudaError_t cudaDeviceCanAccessPeer(int *can, int, int) {
	*can = 1;
	return cudaSuccess;
}
cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
cudaError_t cudaMemcpyPeerAsync(void *d, int, const void *s, int, size_t n, cudaStream_t) {
	memmove(d, s, n);
	return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) {
	memset(d, v, n);
	return cudaSuccess;
}
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) {
	*s = new cusimStream{0};
	return cudaSuccess;
}
cudaError_t cudaStreamDestroy(cudaStream_t s) {
	delete s;
	return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e) {
	*e = new cusimEvent();
	return cudaSuccess;
}
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) {
	delete e;
	return cudaSuccess;
}
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) {
	e->t = std::chrono::steady_clock::now(), e->recorded = true;
	e->busy_polls = cusim::env_int("CUSIM_EVENT_BUSY", 0);
	return cudaSuccess;
}
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventQuery(cudaEvent_t e) {
	if (e->busy_polls > 0) {
		--e->busy_polls;
		return cudaErrorNotReady;
	}
	return cudaSuccess;
}
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) {
	if (!a->recorded || !b->recorded)
		return cudaErrorInvalidValue;
	*ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
	return cudaSuccess;
}
cudaError_t cudaGetDriverEntryPoint(const char *name, void **fn, unsigned long long, cudaDriverEntryPointQueryResult *q) {
	if (strcmp(name, "cuTensorMapEncodeTiled") == 0) {
		*fn = (void *)&cuTensorMapEncodeTiled;
		if (q)
			*q = cudaDriverEntryPointSuccess;
	} else {
		*fn = nullptr;
		if (q)
			*q = cudaDriverEntryPointSymbolNotFound;
	}
	return cudaSuccess;
}

CUresult cuTensorMapEncodeTiled(CUtensorMap *out, CUtensorMapDataType type, cuuint32_t rank, void *global, const cuuint64_t *dims, const cuuint64_t *strides,
                                const cuuint32_t *box, const cuuint32_t *elem_strides, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                CUtensorMapFloatOOBfill) {
	const uint32_t elem = type == CU_TENSOR_MAP_DATA_TYPE_UINT8 ? 1 : type == CU_TENSOR_MAP_DATA_TYPE_UINT16 ? 2 : 4;
	if (rank != 3 || ((size_t)global & 15))
		return CUDA_ERROR_INVALID_VALUE;
	cusim::TensorMap m{};
	m.magic = cusim::kMapMagic, m.base = (uint8_t *)global, m.elem = elem, m.rank = rank;
	for (uint32_t i = 0; i < rank; ++i) {
		if (dims[i] == 0 || box[i] == 0 || box[i] > 256 || elem_strides[i] != 1)
			return CUDA_ERROR_INVALID_VALUE;
		m.dims[i] = dims[i], m.box[i] = box[i];
		m.strides[i] = i == 0 ? elem : strides[i - 1];
		if (i > 0 && (strides[i - 1] & 15)) // global strides are multiples of 16 bytes
			return CUDA_ERROR_INVALID_VALUE;
	}
	if (((size_t)box[0] * elem) & 15) // the innermost box extent is a multiple of 16 bytes
		return CUDA_ERROR_INVALID_VALUE;
	memset(out, 0, sizeof(*out));
	memcpy(out, &m, sizeof(m));
	return CUDA_SUCCESS;
}

} // extern "C"
