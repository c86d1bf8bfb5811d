// tests/cusim -- TEST INFRASTRUCTURE, not product code.
//
// A minimal CPU emulation of the CUDA execution model, just large enough to run hypercraft_b200/csrc/*.cu unchanged on a
// machine without a GPU: every CUDA thread is a fibre with its own stack, a CTA is 1..1024 fibres on one OS thread, warp
// collectives (__shfl_sync, __ballot_sync, __reduce_*_sync, __match_any_sync ...) and __syncthreads are rendez-vous points of
// the fibre scheduler, shared memory is per-CTA host memory, mbarrier + TMA (cp.async.bulk[.tensor]) complete at issue.
// What it proves: the kernels' ARITHMETIC, indexing, barrier structure (a missing barrier shows up as stale data under the
// extreme warp-at-a-time schedule, and the schedule can be reversed), alignment of vector accesses (x86 aligned moves /
// -fsanitize=alignment) and the host-side call sequences of hcm_api.cu.  What it cannot prove: timing, memory-model races
// between truly concurrent warps, sm_100a resource limits.  The CUDA build is the product; `pytest -m gpu` on a B200 is the gate.
//
// Only tests/ builds and loads the result (tests/cusim/_build/libhcm_sim.so); hypercraft_b200/ never does.
#pragma once
#define HCM_CUSIM 1
#include <algorithm>
#include <atomic>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>

// ---- function / variable qualifiers -----------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static thread_local // a CTA runs on one OS thread
#define __constant__ static

// ---- vector types -----------------------------------------------------------------------------------------------------
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(4) char4 { signed char x, y, z, w; };
struct dim3 { unsigned x = 1, y = 1, z = 1; dim3() = default; dim3(unsigned a, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline char4 make_char4(signed char x, signed char y, signed char z, signed char w) { return char4{x, y, z, w}; }

// ---- runtime API subset -----------------------------------------------------------------------------------------------
typedef int cudaError_t;
enum : int { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNotReady = 600, cudaErrorNoDevice = 100 };
struct cusimStream;
struct cusimEvent;
typedef cusimStream *cudaStream_t;
typedef cusimEvent *cudaEvent_t;
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0, cudaEnableDefault = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0, cudaDriverEntryPointSymbolNotFound = 1 };
enum cudaMemcpySrcAccessOrder { cudaMemcpySrcAccessOrderInvalid = 0, cudaMemcpySrcAccessOrderStream = 1 };
struct cudaMemcpyAttributes { cudaMemcpySrcAccessOrder srcAccessOrder; int srcLocHint[2], dstLocHint[2]; unsigned flags; };
struct cudaDeviceProp { char name[256]; int major, minor, multiProcessorCount; size_t totalGlobalMem, sharedMemPerBlockOptin; };

extern "C" {
const char *cudaGetErrorString(cudaError_t);
cudaError_t cudaGetLastError();
cudaError_t cudaGetDeviceCount(int *);
cudaError_t cudaSetDevice(int);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *, int);
cudaError_t cudaMalloc(void **, size_t);
cudaError_t cudaFree(void *);
cudaError_t cudaHostAlloc(void **, size_t, unsigned);
cudaError_t cudaFreeHost(void *);
cudaError_t cudaMemcpy(void *, const void *, size_t, cudaMemcpyKind);
cudaError_t cudaMemcpyAsync(void *, const void *, size_t, cudaMemcpyKind, cudaStream_t);
cudaError_t cudaMemcpyBatchAsync(void **dsts, void **srcs, size_t *sizes, size_t count, cudaMemcpyAttributes *attrs, size_t *attr_idx, size_t n_attrs,
                                 size_t *fail_idx, cudaStream_t);
cudaError_t cudaDeviceCanAccessPeer(int *can, int device, int peer);
cudaError_t cudaDeviceEnablePeerAccess(int peer, unsigned flags);
cudaError_t cudaMemcpyPeerAsync(void *dst, int dst_device, const void *src, int src_device, size_t, cudaStream_t);
cudaError_t cudaMemsetAsync(void *, int, size_t, cudaStream_t);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *, unsigned);
cudaError_t cudaStreamDestroy(cudaStream_t);
cudaError_t cudaStreamSynchronize(cudaStream_t);
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned);
cudaError_t cudaEventCreate(cudaEvent_t *);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *, unsigned);
cudaError_t cudaEventDestroy(cudaEvent_t);
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t);
cudaError_t cudaEventSynchronize(cudaEvent_t);
cudaError_t cudaEventQuery(cudaEvent_t);
cudaError_t cudaEventElapsedTime(float *, cudaEvent_t, cudaEvent_t);
cudaError_t cudaGetDriverEntryPoint(const char *, void **, unsigned long long, cudaDriverEntryPointQueryResult *);
}
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
template <typename T> static inline cudaError_t cudaMemcpyFromSymbol(void *dst, const T &symbol, size_t n) { return memcpy(dst, &symbol, n), cudaSuccess; }
template <typename T> static inline cudaError_t cudaMemcpyToSymbol(T &symbol, const void *src, size_t n) { return memcpy(&symbol, src, n), cudaSuccess; }

// ---- the fibre machine --------------------------------------------------------------------------------------------------
namespace cusim {
struct Idx { unsigned x, y, z; };
struct ThreadState { // what a CUDA thread can ask about itself
	Idx tid, bid, bdim, gdim;
	unsigned lane, warp;
};
ThreadState &self();
unsigned char *dynamic_smem();
void launch(dim3 grid, dim3 block, size_t smem, std::function<void()> body); // synchronous
void sync_threads();
int sync_threads_or(int);
int sync_threads_and(int);
void yield(); // a spin-wait iteration
// warp rendez-vous: every lane named in `mask` deposits `v`; returns a pointer to the 32 deposited values (valid until the
// lane's next collective with the same mask)
const unsigned long long *warp_gather(unsigned mask, unsigned long long v);
// mbarrier / TMA
void mbar_init(unsigned long long *bar, unsigned count);
void mbar_expect_tx(unsigned long long *bar, unsigned bytes);
void mbar_wait(unsigned long long *bar, unsigned parity);
void tma_load_1d(void *dst, const void *src, unsigned bytes, unsigned long long *bar);
void tma_load_3d(void *dst, const void *map, int c0, int c1, int c2, unsigned long long *bar);
} // namespace cusim

#define threadIdx (::cusim::self().tid)
#define blockIdx (::cusim::self().bid)
#define blockDim (::cusim::self().bdim)
#define gridDim (::cusim::self().gdim)

#define HCM_LAUNCH(kernel, grid, block, smem, stream, ...) ::cusim::launch((grid), (block), (smem), [=]() { kernel(__VA_ARGS__); })
#define HCM_DYNAMIC_SMEM(name) uint8_t *name = ::cusim::dynamic_smem()

// ---- device intrinsics --------------------------------------------------------------------------------------------------
static inline void __syncthreads() { ::cusim::sync_threads(); }
static inline int __syncthreads_or(int p) { return ::cusim::sync_threads_or(p); }
static inline int __syncthreads_and(int p) { return ::cusim::sync_threads_and(p); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { ::cusim::warp_gather(mask, 0); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) { return (unsigned)(((((unsigned long long)hi) << 32) | lo) >> (s & 31)); }
static inline unsigned __funnelshift_lc(unsigned lo, unsigned hi, unsigned s) {
	s = s > 32 ? 32 : s;
	return (unsigned)((((((unsigned long long)hi) << 32) | lo) << s) >> 32);
}
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel) {
	const unsigned long long v = ((unsigned long long)b << 32) | a;
	unsigned r = 0;
	for (int i = 0; i < 4; ++i)
		r |= (unsigned)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xff) << (8 * i);
	return r;
}
template <typename T> static inline T __ldg(const T *p) { return *p; }
template <typename T> static inline T __ldcg(const T *p) { return *p; }
static inline size_t __cvta_generic_to_shared(const void *p) { return (size_t)p; }
static inline long long clock64() { return 0; }

template <typename T> static inline T __shfl_sync(unsigned mask, T v, int src) {
	static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
	unsigned long long raw = 0;
	memcpy(&raw, &v, sizeof(T));
	const unsigned long long *s = ::cusim::warp_gather(mask, raw);
	T out;
	memcpy(&out, &s[src & 31], sizeof(T));
	return out;
}
template <typename T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta) {
	unsigned long long raw = 0;
	memcpy(&raw, &v, sizeof(T));
	const unsigned long long *s = ::cusim::warp_gather(mask, raw);
	const unsigned lane = ::cusim::self().lane;
	T out;
	memcpy(&out, &s[lane >= delta ? lane - delta : lane], sizeof(T));
	return out;
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
	const unsigned long long *s = ::cusim::warp_gather(mask, pred ? 1 : 0);
	unsigned r = 0;
	for (int i = 0; i < 32; ++i)
		if (((mask >> i) & 1) && s[i])
			r |= 1u << i;
	return r;
}
static inline unsigned __match_any_sync(unsigned mask, unsigned v) {
	const unsigned long long *s = ::cusim::warp_gather(mask, v);
	unsigned r = 0;
	for (int i = 0; i < 32; ++i)
		if (((mask >> i) & 1) && (unsigned)s[i] == v)
			r |= 1u << i;
	return r;
}
#define CUSIM_REDUCE(name, init, op)                                                                                   \
	static inline unsigned name(unsigned mask, unsigned v) {                                                           \
		const unsigned long long *s = ::cusim::warp_gather(mask, v);                                                   \
		unsigned r = init;                                                                                             \
		for (int i = 0; i < 32; ++i)                                                                                   \
			if ((mask >> i) & 1) {                                                                                     \
				const unsigned x = (unsigned)s[i];                                                                     \
				r = op;                                                                                                \
			}                                                                                                          \
		return r;                                                                                                      \
	}
CUSIM_REDUCE(__reduce_or_sync, 0u, r | x)
CUSIM_REDUCE(__reduce_and_sync, 0xffffffffu, r &x)
CUSIM_REDUCE(__reduce_add_sync, 0u, r + x)
CUSIM_REDUCE(__reduce_min_sync, 0xffffffffu, (x < r ? x : r))
CUSIM_REDUCE(__reduce_max_sync, 0u, (x > r ? x : r))
#undef CUSIM_REDUCE

// atomics (global or shared: both are host memory; CTAs of one launch run on several OS threads)
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicOr(unsigned *p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicAnd(unsigned *p, unsigned v) { return __atomic_fetch_and(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicExch(unsigned *p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicMin(unsigned *p, unsigned v) {
	unsigned o = __atomic_load_n(p, __ATOMIC_RELAXED);
	while (v < o && !__atomic_compare_exchange_n(p, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
	}
	return o;
}
static inline unsigned atomicMax(unsigned *p, unsigned v) {
	unsigned o = __atomic_load_n(p, __ATOMIC_RELAXED);
	while (v > o && !__atomic_compare_exchange_n(p, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
	}
	return o;
}
static inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v) {
	unsigned long long o = __atomic_load_n(p, __ATOMIC_RELAXED);
	while (v > o && !__atomic_compare_exchange_n(p, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
	}
	return o;
}

// CUDA's overload set of min / max (mixed signedness promotes to unsigned)
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, int b) { return min(a, (unsigned)b); }
static inline unsigned min(int a, unsigned b) { return min((unsigned)a, b); }
static inline unsigned max(unsigned a, int b) { return max(a, (unsigned)b); }
static inline unsigned max(int a, unsigned b) { return max((unsigned)a, b); }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
static inline unsigned long min(unsigned long a, unsigned long b) { return a < b ? a : b; }
static inline unsigned long max(unsigned long a, unsigned long b) { return a > b ? a : b; }
