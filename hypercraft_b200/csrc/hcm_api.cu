// C ABI of the B200 chunk mesher (include/hcm.h): context, residency, batching, output arenas.
// All device work is enqueued on the ctx stream; there is no CPU fallback anywhere in this file.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include <immintrin.h>

#include "hcm_device.cuh"

namespace hcm {
void launch_derive(const DevWorld &w, const UploadEntry *entries, uint32_t n, const uint8_t *prop_index, const DevTables *tables,
                   const int8_t *raw_heights, int variant, bool trial, cudaStream_t st);
void launch_derive_digest(const DevWorld &w, const UploadEntry *entries, uint32_t n, const int8_t *raw_heights, const unsigned long long *expect,
                          unsigned long long *out, uint32_t *mismatches, cudaStream_t st);
void launch_mesh_digest(const hcm_task_result *task_out, const hcm_mesh_desc *meshes, const unsigned long long *vtx, const uint16_t *idx, uint32_t n,
                        uint32_t mesh_cap, unsigned long long vtx_cap, unsigned long long idx_cap, const unsigned long long *expect, unsigned long long *out,
                        uint32_t *mismatches, cudaStream_t st);
void launch_clear(const DevWorld &w, const UploadEntry *entries, uint32_t n, cudaStream_t st);
void launch_flood_slots(uint4 *out, cudaStream_t st);
void launch_edit(const DevWorld &w, const hcm_block_edit *edits, uint32_t n, const uint2 *columns, uint32_t n_columns, uint32_t *result, const uint8_t *prop_index,
                 const DevTables *tables, int8_t *raw_heights, hcm_sunlight_change *changes, uint32_t *n_changes, uint32_t cap, cudaStream_t st);
void launch_classify(const DevWorld &w, const TaskIn *tasks, uint32_t n, uint32_t *worklist, BatchHeads *heads, hcm_task_result *task_out, int force_all,
                     cudaStream_t st);
int derive_variant_count();
int mesh_variant_count();
cudaError_t mesh_kernel_prepare();
int mesh_kernel_smem_bytes();
void launch_mesh(const MeshMaps &maps, const MeshParams &p, int grid, int variant, cudaStream_t st);
void launch_scatter(const hcm_mesh_placement *placements, uint32_t n, const hcm_mesh_desc *meshes, const unsigned long long *vtx, const uint16_t *idx, void *vdst,
                    void *idst, int grid, cudaStream_t st);
} // namespace hcm

using namespace hcm;

static thread_local std::string g_create_error;

// Persistent host worker threads (the uniform-chunk scan of large uploads): spawning 15 threads per call costs ~0.5 ms.
struct WorkerPool {
	std::vector<std::thread> threads;
	std::mutex mu;
	std::condition_variable cv, cv_done;
	std::function<void()> job;
	uint64_t generation = 0;
	uint32_t running = 0;
	bool stop = false;
	void start(uint32_t n) {
		for (uint32_t i = 0; i < n; ++i)
			threads.emplace_back([this]() {
				uint64_t seen = 0;
				for (;;) {
					std::function<void()> f;
					{
						std::unique_lock<std::mutex> lk(mu);
						cv.wait(lk, [&] { return stop || generation != seen; });
						if (stop)
							return;
						seen = generation, f = job;
					}
					f();
					{
						std::lock_guard<std::mutex> lk(mu);
						if (--running == 0)
							cv_done.notify_all();
					}
				}
			});
	}
	void launch(std::function<void()> f) {
		std::lock_guard<std::mutex> lk(mu);
		job = std::move(f), ++generation, running = (uint32_t)threads.size();
		cv.notify_all();
	}
	void wait() {
		std::unique_lock<std::mutex> lk(mu);
		cv_done.wait(lk, [&] { return running == 0; });
	}
	~WorkerPool() {
		{
			std::lock_guard<std::mutex> lk(mu);
			stop = true;
		}
		cv.notify_all();
		for (auto &t : threads)
			t.join();
	}
};

// Pipeline stages ("blocks") of one upload: stage 0 holds `first` chunks, every later stage c->upload_block.
struct BlockLayout {
	uint32_t n, first, size;
	uint32_t count() const { return n <= first ? 1u : 1u + (n - first + size - 1) / size; }
	uint32_t of(uint32_t i) const { return i < first ? 0u : 1u + (i - first) / size; }
	uint32_t lo(uint32_t b) const { return b == 0 ? 0u : std::min(n, first + (b - 1) * size); }
	uint32_t hi(uint32_t b) const { return std::min(n, first + b * size); }
};

// Qualification of the kernel variants on the device (hcm.h, "Kernel variants")
struct KernelTune {
	int forced = -1; // 0 .. n: environment / hcm_set_option; -1: the A/B rounds decide
	uint32_t n_variants = 0;
	uint32_t selected = 0, state = HCM_TUNE_PENDING, rounds = 0;
	uint32_t mismatches[HCM_MAX_VARIANTS] = {0, 0, 0};
	float sum_base = 0.f, sum_var[HCM_MAX_VARIANTS] = {0.f, 0.f, 0.f};
	bool in_flight = false; // an A/B round is in the stream; its verdict is read at the next launch or host synchronisation
	bool invalid = false;   // ... and must be discarded
};
constexpr int kTuneRuns = HCM_MAX_VARIANTS + 2; // baseline, the variants, baseline again
struct TuneHost { // pinned
	uint32_t mismatch[2][HCM_MAX_VARIANTS]; // [derive, mesh][variant - 1]
	BatchHeads heads[kTuneRuns];             // allocation heads after each run of a mesh round (an overflowed run proves nothing)
};

struct hcm_ctx {
	bool tune = true; // env HCM_TUNE
	bool tune_fault = false; // test knob: corrupt the variant's hash of unit 0 before the comparison (the rejection path)
	uint32_t tune_min_chunks = 2048, tune_min_tasks = 2048, tune_rounds = 2;
	KernelTune tune_derive, tune_mesh;
	unsigned long long *d_digest[2] = {nullptr, nullptr}; // per chunk / per task hashes of an A/B round: [0] the baseline's output, [1] the variant under comparison
	uint32_t *d_tune_mismatch = nullptr;                   // [2]
	TuneHost *h_tune = nullptr;
	bool batch_ab = false; // the last mesh batch was an A/B round (which events bracket its classify kernel)
	BlockLayout layout{0, 1, 1}; // of the upload in progress (hcm_upload_and_mesh buckets its tasks by stage)
	int device = 0;
	cudaStream_t stream = nullptr;
	cudaStream_t copy_stream = nullptr; // host->device copies of large uploads overlap the host-side scan and the derive kernel
	std::vector<cudaEvent_t> block_events;
	cudaStream_t fetch_stream = nullptr; // device->host copies of hcm_upload_and_mesh overlap uploads and kernels
	std::vector<cudaEvent_t> sub_events;
	BatchHeads *h_heads_ring = nullptr; // pinned: heads after each sub-batch
	uint32_t heads_ring_cap = 0;
	hcm_task_result *h_task_stage = nullptr; // pinned: task results in sub-batch order
	uint64_t h_task_stage_cap = 0;
	std::vector<int32_t> cell_block; // scratch of hcm_upload_and_mesh
	// hcm_set_blocks
	hcm_block_edit *d_edits = nullptr;
	uint2 *d_edit_columns = nullptr;
	uint32_t *d_edit_result = nullptr; // [cap] result words, then the change counter
	hcm_sunlight_change *d_sun_changes = nullptr;
	uint32_t edit_cap = 0, sun_cap = 0;
	std::vector<hcm_sunlight_change> edit_sunlight;
	std::vector<int16_t> edit_remesh;
	std::vector<uint8_t> edit_remesh_flags;
	bool use_batch_copy = true;
	// pipeline granularity (env HCM_UPLOAD_BLOCK, HCM_MIN_SUB_BATCH: test knobs): chunks per pipeline stage (2048 = 128 MB of Block
	// data); ready sets smaller than min_sub_batch wait for the next upload block
	uint32_t upload_block = 2048, min_sub_batch = 512;
	// first pipeline stage of a threaded upload (env HCM_UPLOAD_FIRST_BLOCK): small, because it usually crosses the bus unscanned
	uint32_t first_block = 256;
	bool adaptive_upload = true; // env HCM_UPLOAD_ADAPTIVE=0: always wait for the scan (the round-1 behaviour)
	hcm_mesh_placement *d_placements = nullptr; // hcm_mesh_emit_into
	uint32_t placement_cap = 0;
	uint32_t sink_meshes = 0;
	bool counts_valid = false; // h_task_out / h_mesh_out hold the last batch's rows (hcm_mesh_counts, hcm_mesh_fetch, hcm_upload_and_mesh)
	std::unique_ptr<WorkerPool> pool;
	uint32_t host_threads = 0; // scan threads of hcm_upload_chunks (env HCM_HOST_THREADS; 0 = hardware_concurrency - 1, at most 15)
	DevWorld w{};
	size_t n_cells = 0; // window chunks
	uint32_t max_chunks = 0, max_tasks = 0;

	DevTables *d_tables = nullptr;
	uint8_t *d_prop_index = nullptr;
	hcm_custom_mesh *d_meshes = nullptr;
	DevMeshInfo *d_mesh_info = nullptr;
	uint16_t *d_raw = nullptr;
	int8_t *d_rawh = nullptr;
	UploadEntry *d_entries = nullptr;
	MeshMaps maps{}; // TMA descriptors: property grid (halo box), indirect-light-pass planes and sunlight heights (flood inputs)

	std::vector<int32_t> slot_of; // host mirror of DevWorld::slot
	std::vector<uint8_t> nbr_count; // per cell: resident chunks in its 3x3x3 neighbourhood (27 = a mesh task is possible)
	uint64_t residency_version = 1, all_entries_version = 0; // cache of the "every resident chunk" entry list on the device
	UploadEntry *d_all_entries = nullptr;
	uint32_t n_all_entries = 0;
	std::vector<int32_t> free_slots;

	TaskIn *d_tasks = nullptr;
	uint32_t *d_worklist = nullptr;
	BatchHeads *d_heads = nullptr;
	hcm_task_result *d_task_out = nullptr;
	hcm_mesh_desc *d_mesh_out = nullptr;
	uint64_t *d_vtx = nullptr;
	uint16_t *d_idx = nullptr;
	uint64_t vtx_cap = 0, idx_cap = 0;
	uint32_t mesh_cap = 0;
	DynRec *d_dyn = nullptr;
	uint4 *d_flood_slots = nullptr; // per-thread flood columns of the mesh kernel
	uint8_t *d_debug_light = nullptr;
	int grid = 0;

	BatchHeads *h_heads = nullptr; // pinned
	hcm_task_result *h_task_out = nullptr;
	hcm_mesh_desc *h_mesh_out = nullptr;
	uint64_t *h_vtx = nullptr;
	uint16_t *h_idx = nullptr;
	uint64_t h_vtx_cap = 0, h_idx_cap = 0;
	uint32_t h_mesh_cap = 0, h_task_cap = 0;

	std::vector<TaskIn> tasks;
	uint32_t batch_n = 0;
	bool batch_pending = false, batch_done = false;
	int force_all = 0;

	cudaEvent_t ev[12]{};
	cudaEvent_t ev_tune[2][2 * kTuneRuns + 1]{}; // [derive, mesh]: begin / end of every run of an A/B round, then "verdict on the host"
	cudaEvent_t ev_peer = nullptr; // orders hcm_copy_chunks_from between two contexts' streams
	// task lists go up from two pinned buffers in turn, so that hcm_mesh_batch never waits for the stream (a pageable source makes
	// cudaMemcpyAsync wait for the previous batch's kernels) and batches can be enqueued back to back
	TaskIn *h_task_ring[2] = {nullptr, nullptr};
	cudaEvent_t task_ring_ev[2] = {nullptr, nullptr};
	bool task_ring_used[2] = {false, false};
	int task_ring_next = 0;
	uint64_t peers_tried = 0;      // bit d: direct access to device d has been asked for (hcm_copy_chunks_from)
	bool ev_upload = false, ev_batch = false, ev_fetch = false, ev_sink = false;
	hcm_stats stats{};
	std::string err;
};

namespace {

enum { EV_UP0, EV_UP1, EV_UP2, EV_B0, EV_B1, EV_B2, EV_F0, EV_F1, EV_S0, EV_S1,
       EV_BC, // end of the classify kernel in a mesh A/B round (EV_B1 .. EV_B2 bracket the round's last baseline run there)
       EV_COUNT };
static_assert(EV_COUNT <= 12, "hcm_ctx::ev");
constexpr int kTuneDone = 2 * kTuneRuns;

int fail(hcm_ctx *c, int status, const char *fmt, ...) {
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	if (c)
		c->err = buf;
	else
		g_create_error = buf;
	return status;
}

#define CU(call)                                                                                                       \
	do {                                                                                                               \
		cudaError_t e_ = (call);                                                                                       \
		if (e_ != cudaSuccess)                                                                                         \
			return fail(c, HCM_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
	} while (0)

template <typename T> cudaError_t dev_alloc(hcm_ctx *c, T **p, size_t count) {
	cudaError_t e = cudaMalloc((void **)p, count * sizeof(T));
	if (e == cudaSuccess)
		c->stats.device_bytes += count * sizeof(T);
	return e;
}

bool in_window(const hcm_ctx *c, const int16_t *pos, int *cx, int *cy, int *cz) {
	*cx = pos[0] - c->w.ox, *cy = pos[1] - c->w.oy, *cz = pos[2] - c->w.oz;
	return *cx >= 0 && *cx < c->w.wx && *cy >= 0 && *cy < c->w.wy && *cz >= 0 && *cz < c->w.wz;
}
size_t cell(const hcm_ctx *c, int cx, int cy, int cz) { return ((size_t)cy * c->w.wz + cz) * c->w.wx + cx; }

// true when all 32768 voxels hold the same Block value (period-4 array whose first four elements are equal)
inline bool chunk_is_uniform_generic(const uint16_t *b) {
	const uint16_t v = b[0];
	return b[1] == v && b[2] == v && b[3] == v && memcmp(b, b + 4, (HCM_CHUNK_VOXELS - 4) * 2) == 0;
}
// one aligned-or-not 32-byte load per 16 voxels, early exit every 256 bytes (terrain chunks differ within the first rows).
// The scan of a uniform chunk is a pure DRAM stream over 4 KB pages, where the hardware prefetchers stop at every page
// boundary: software prefetches 4 KB ahead (into the next chunk at the end; a prefetch never faults) measured +10..20 % per
// thread and at 8 threads on the build host (11 -> 12.5 GB/s, 78 -> 88 GB/s).
constexpr int kScanPrefetchAhead = 4096;
__attribute__((target("avx2"))) bool chunk_is_uniform_avx2(const uint16_t *b) {
	const __m256i v = _mm256_set1_epi16((short)b[0]);
	const __m256i *p = reinterpret_cast<const __m256i *>(b);
	const char *ahead = reinterpret_cast<const char *>(b) + kScanPrefetchAhead;
	for (int i = 0; i < HCM_CHUNK_VOXELS / 16; i += 8) {
		_mm_prefetch(ahead + i * 32, _MM_HINT_T0), _mm_prefetch(ahead + i * 32 + 64, _MM_HINT_T0);
		_mm_prefetch(ahead + i * 32 + 128, _MM_HINT_T0), _mm_prefetch(ahead + i * 32 + 192, _MM_HINT_T0);
		__m256i a0 = _mm256_xor_si256(_mm256_loadu_si256(p + i), v), a1 = _mm256_xor_si256(_mm256_loadu_si256(p + i + 1), v);
		__m256i a2 = _mm256_xor_si256(_mm256_loadu_si256(p + i + 2), v), a3 = _mm256_xor_si256(_mm256_loadu_si256(p + i + 3), v);
		a0 = _mm256_or_si256(a0, _mm256_xor_si256(_mm256_loadu_si256(p + i + 4), v)), a1 = _mm256_or_si256(a1, _mm256_xor_si256(_mm256_loadu_si256(p + i + 5), v));
		a2 = _mm256_or_si256(a2, _mm256_xor_si256(_mm256_loadu_si256(p + i + 6), v)), a3 = _mm256_or_si256(a3, _mm256_xor_si256(_mm256_loadu_si256(p + i + 7), v));
		const __m256i a = _mm256_or_si256(_mm256_or_si256(a0, a1), _mm256_or_si256(a2, a3));
		if (!_mm256_testz_si256(a, a))
			return false;
	}
	return true;
}
const bool g_have_avx2 = __builtin_cpu_supports("avx2");
inline bool chunk_is_uniform(const uint16_t *b) { return g_have_avx2 ? chunk_is_uniform_avx2(b) : chunk_is_uniform_generic(b); }

constexpr uint32_t kScanGrain = 32;     // chunks per work item of the scan threads

// a chunk became resident (+1) or was removed (-1): update the 27 cells that see it as a neighbour
void bump_neighbours(hcm_ctx *c, int cx, int cy, int cz, int delta) {
	for (int dy = -1; dy <= 1; ++dy)
		for (int dz = -1; dz <= 1; ++dz)
			for (int dx = -1; dx <= 1; ++dx) {
				const int x = cx + dx, y = cy + dy, z = cz + dz;
				if (x >= 0 && x < c->w.wx && y >= 0 && y < c->w.wy && z >= 0 && z < c->w.wz)
					c->nbr_count[cell(c, x, y, z)] = (uint8_t)(c->nbr_count[cell(c, x, y, z)] + delta);
			}
	++c->residency_version;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                    const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int build_tables(hcm_ctx *c, const hcm_registry *reg, DevTables &t, std::vector<DevMeshInfo> &minfo) {
	memset(&t, 0, sizeof(t));
	if (!reg || !reg->prop_index || !reg->props || !reg->tex_flags || (reg->n_meshes && !reg->meshes))
		return fail(c, HCM_ERR_INVALID, "registry has null tables");
	if (reg->n_props == 0 || reg->n_props > 256 || reg->n_meshes > 255)
		return fail(c, HCM_ERR_INVALID, "registry sizes out of range (props %u, meshes %u)", reg->n_props, reg->n_meshes);
	for (uint32_t i = 0; i < 65536; ++i)
		if (reg->prop_index[i] >= reg->n_props)
			return fail(c, HCM_ERR_INVALID, "prop_index[%u] = %u out of range", i, reg->prop_index[i]);
	// texture codes: one per distinct BlockTexture value (equality on the whole u16, BlockTexture.hpp:78-79)
	std::map<uint16_t, uint32_t> codes;
	codes[0] = 0;
	auto code_of = [&](uint16_t tex) -> int {
		auto it = codes.find(tex);
		if (it != codes.end())
			return (int)it->second;
		if (codes.size() >= 256)
			return -1;
		uint32_t k = (uint32_t)codes.size();
		codes[tex] = k;
		return (int)k;
	};
	auto tex_flags = [&](uint16_t tex, uint8_t *out) -> bool {
		uint32_t id = tex & 0x1fffu;
		if (id >= reg->n_textures)
			return false;
		*out = (uint8_t)((reg->tex_flags[id] & 7u) | (id == 0 ? TF_EMPTY : 0));
		return true;
	};
	minfo.assign(reg->n_meshes, DevMeshInfo{});
	for (uint32_t m = 0; m < reg->n_meshes; ++m) {
		const hcm_custom_mesh &cm = reg->meshes[m];
		DevMeshInfo &mi = minfo[m];
		if (cm.face_count > HCM_MAX_MESH_FACES)
			return fail(c, HCM_ERR_INVALID, "custom mesh %u has %u faces", m, cm.face_count);
		if (cm.dynamic_kind != HCM_DYNAMIC_NONE && cm.dynamic_kind != HCM_DYNAMIC_LIQUID_FLOW)
			return fail(c, HCM_ERR_UNSUPPORTED, "custom mesh %u uses unknown dynamic generator kind %u", m, cm.dynamic_kind);
		mi.dynamic_kind = (uint8_t)cm.dynamic_kind;
		for (uint32_t i = 0; i < cm.face_count; ++i) {
			uint8_t fl;
			if (!tex_flags(cm.faces[i].texture, &fl))
				return fail(c, HCM_ERR_INVALID, "custom mesh %u face %u: texture id out of range", m, i);
			if (cm.faces[i].axis > 2 || cm.faces[i].light_face > 5 || cm.faces[i].render_face > 5)
				return fail(c, HCM_ERR_INVALID, "custom mesh %u face %u: bad axis/face", m, i);
			if (fl & HCM_TEX_TRANS_PASS)
				mi.n_trans++, mi.trans_mask |= (uint16_t)(1u << i);
			else
				mi.n_opaque++;
		}
		if (cm.dynamic_kind == HCM_DYNAMIC_LIQUID_FLOW) {
			uint8_t fl;
			int code = code_of(cm.liquid_flow_tex);
			if (code < 0 || !tex_flags(cm.liquid_flow_tex, &fl))
				return fail(c, HCM_ERR_UNSUPPORTED, "too many distinct textures / bad liquid texture");
			mi.flow_code = (uint8_t)code;
			mi.flow_trans = (fl & HCM_TEX_TRANS_PASS) ? 1 : 0;
			mi.height_unit = cm.liquid_height_unit;
			mi.variant = cm.liquid_variant;
		}
	}
	std::vector<uint8_t> levels;
	for (uint32_t i = 0; i < reg->n_props; ++i) {
		const hcm_block_prop &p = reg->props[i];
		if (p.transparency > 2 || p.collision > 2 || p.light_level > 15 || p.custom_mesh > reg->n_meshes)
			return fail(c, HCM_ERR_INVALID, "property row %u out of range", i);
		bool all_solid = true, all_empty = true;
		for (int f = 0; f < 6; ++f) {
			int code = code_of(p.tex[f]);
			uint8_t fl;
			if (code < 0)
				return fail(c, HCM_ERR_UNSUPPORTED, "more than 255 distinct block textures");
			if (!tex_flags(p.tex[f], &fl))
				return fail(c, HCM_ERR_INVALID, "property row %u: texture id out of range", i);
			t.ptex[i][f] = (uint8_t)code;
			if (fl & TF_EMPTY)
				all_solid = false;
			else
				all_empty = false;
			if (fl & (HCM_TEX_TRANSPARENT | HCM_TEX_LIQUID))
				all_solid = false;
		}
		const bool dynamic = p.custom_mesh && reg->meshes[p.custom_mesh - 1].dynamic_kind != HCM_DYNAMIC_NONE;
		uint8_t fl = 0;
		if (p.transparency != 0)
			fl |= PF_INDIRECT;
		if (p.transparency == 2 || p.collision != 1)
			fl |= PF_DIRECT;
		if (all_solid && !p.custom_mesh)
			fl |= PF_SOLID;
		if (all_empty && !dynamic)
			fl |= PF_EMPTY;
		if (p.custom_mesh)
			fl |= PF_CUSTOM;
		if (dynamic)
			fl |= PF_DYNAMIC;
		if (p.liquid_role == HCM_LIQUID_SOURCE)
			fl |= PF_LIQ_SOURCE;
		if (p.liquid_role == HCM_LIQUID_FLOW)
			fl |= PF_LIQ_FLOW;
		t.pflags[i] = fl;
		t.pmesh[i] = p.custom_mesh;
		bool one_texture = !dynamic;
		for (int f = 1; f < 6; ++f)
			one_texture = one_texture && p.tex[f] == p.tex[0];
		t.pmisc[i] = (uint8_t)((p.liquid_obstacle ? 1 : 0) | (p.transparency == 2 ? 2 : 0) | (one_texture ? 4 : 0) |
		                       (p.light_level << 4)); // bit1: GetVerticalLightPass, Block.hpp:70
		if (p.light_level && std::find(levels.begin(), levels.end(), p.light_level) == levels.end())
			levels.push_back(p.light_level);
	}
	for (const auto &kv : codes) {
		t.tex16[kv.second] = kv.first;
		uint8_t fl = 0;
		tex_flags(kv.first, &fl);
		t.texflags[kv.second] = fl;
	}
	if (levels.size() > (size_t)kMaxLightLevels)
		return fail(c, HCM_ERR_UNSUPPORTED, "registry has %zu distinct light levels, at most %d supported", levels.size(), kMaxLightLevels);
	std::sort(levels.begin(), levels.end(), [](uint8_t a, uint8_t b) { return a > b; });
	c->w.n_levels = (int)levels.size();
	for (size_t i = 0; i < levels.size(); ++i)
		c->w.levels[i] = levels[i];
	return HCM_OK;
}

int alloc_arenas(hcm_ctx *c, uint64_t vtx_cap, uint32_t mesh_cap) {
	if (c->d_vtx) {
		c->stats.device_bytes -= c->vtx_cap * 8 + c->idx_cap * 2;
		cudaFree(c->d_vtx), cudaFree(c->d_idx);
		c->d_vtx = nullptr, c->d_idx = nullptr;
	}
	if (c->d_mesh_out && mesh_cap != c->mesh_cap) {
		c->stats.device_bytes -= (size_t)c->mesh_cap * sizeof(hcm_mesh_desc);
		cudaFree(c->d_mesh_out);
		c->d_mesh_out = nullptr;
	}
	vtx_cap = (vtx_cap + 3) & ~3ull;
	c->vtx_cap = vtx_cap, c->idx_cap = vtx_cap / 4 * 6;
	CU(dev_alloc(c, &c->d_vtx, c->vtx_cap));
	CU(dev_alloc(c, &c->d_idx, c->idx_cap));
	if (!c->d_mesh_out) {
		c->mesh_cap = mesh_cap;
		CU(dev_alloc(c, &c->d_mesh_out, c->mesh_cap));
	}
	return HCM_OK;
}

template <typename T> int host_reserve(hcm_ctx *c, T **p, uint64_t *cap, uint64_t need) {
	if (need <= *cap)
		return HCM_OK;
	if (*p)
		cudaFreeHost(*p);
	*p = nullptr;
	uint64_t ncap = std::max<uint64_t>(need + need / 4, 1024);
	CU(cudaHostAlloc((void **)p, ncap * sizeof(T), cudaHostAllocDefault));
	*cap = ncap;
	return HCM_OK;
}

MeshParams mesh_params(const hcm_ctx *c, uint32_t first_task) {
	MeshParams p{};
	p.w = c->w;
	p.tables = c->d_tables, p.meshes = c->d_meshes, p.mesh_info = c->d_mesh_info;
	p.tasks = c->d_tasks + first_task, p.worklist = c->d_worklist, p.heads = c->d_heads;
	p.task_out = c->d_task_out + first_task, p.mesh_out = c->d_mesh_out, p.vtx = c->d_vtx, p.idx = c->d_idx;
	p.vtx_cap = c->vtx_cap, p.idx_cap = c->idx_cap, p.mesh_cap = c->mesh_cap;
	p.dyn_scratch = c->d_dyn, p.debug_light = c->d_debug_light;
	p.flood_slots = c->d_flood_slots;
	return p;
}

// grow a pinned host array, keeping its first `used` elements (copies into it may still be in flight on `st`)
template <typename T> int host_reserve_keep(hcm_ctx *c, T **p, uint64_t *cap, uint64_t need, uint64_t used, cudaStream_t st) {
	if (need <= *cap)
		return HCM_OK;
	T *np_ = nullptr;
	const uint64_t ncap = std::max<uint64_t>({need + need / 4, *cap * 2, (uint64_t)1024});
	CU(cudaHostAlloc((void **)&np_, ncap * sizeof(T), cudaHostAllocDefault));
	if (*p) {
		CU(cudaStreamSynchronize(st));
		if (used)
			memcpy(np_, *p, used * sizeof(T));
		cudaFreeHost(*p);
	}
	*p = np_, *cap = ncap;
	return HCM_OK;
}

// ---- kernel variants: A/B rounds on the device (hcm.h, "Kernel variants") -------------------------------------------------
// Run r of a round: 0 = baseline, 1 .. n = variant r, n + 1 = baseline again (the call's result).
void tune_decide(KernelTune &t, const float *run_ms, const uint32_t *mismatch, uint32_t want_rounds) {
	t.in_flight = false;
	if (t.invalid || t.forced >= 0 || t.state != HCM_TUNE_PENDING)
		return;
	const uint32_t n = t.n_variants;
	// rounds differ in size (an upload block, a whole re-derive): the forms are compared on the SUMS over the same rounds
	t.sum_base += std::min(run_ms[0], run_ms[n + 1]); // the faster of the baseline's two runs: the first one of a small round is cache-cold
	for (uint32_t v = 0; v < n; ++v)
		t.sum_var[v] += run_ms[v + 1], t.mismatches[v] += mismatch[v];
	if (++t.rounds < want_rounds)
		return;
	uint32_t best = 0;
	bool any_bad = false;
	for (uint32_t v = 0; v < n; ++v) {
		if (t.mismatches[v]) {
			any_bad = true; // this device does not reproduce the baseline's bytes with it: never launched again
			continue;
		}
		if (t.sum_var[v] < 0.98f * t.sum_base && (!best || t.sum_var[v] < t.sum_var[best - 1]))
			best = v + 1;
	}
	t.selected = best;
	t.state = best ? HCM_TUNE_VARIANT_ADOPTED : (any_bad ? HCM_TUNE_VARIANT_REJECTED : HCM_TUNE_BASELINE_KEPT);
}

// read the verdict of every round that has completed (never waits)
int tune_harvest(hcm_ctx *c) {
	for (int k = 0; k < 2; ++k) {
		KernelTune &t = k == 0 ? c->tune_derive : c->tune_mesh;
		if (!t.in_flight)
			continue;
		const cudaError_t q = cudaEventQuery(c->ev_tune[k][kTuneDone]);
		if (q == cudaErrorNotReady) {
			(void)cudaGetLastError(); // an answer, not an error
			continue;
		}
		CU(q);
		float ms[kTuneRuns] = {};
		for (uint32_t r = 0; r < t.n_variants + 2; ++r) {
			CU(cudaEventElapsedTime(&ms[r], c->ev_tune[k][2 * r], c->ev_tune[k][2 * r + 1]));
			if (k == 1 && c->h_tune->heads[r].overflow)
				t.invalid = true;
		}
		tune_decide(t, ms, c->h_tune->mismatch[k], c->tune_rounds);
	}
	return HCM_OK;
}

bool tune_wants_round(const hcm_ctx *c, const KernelTune &t, uint32_t units, uint32_t min_units) {
	return c->tune && t.n_variants && t.forced < 0 && t.state == HCM_TUNE_PENDING && !t.in_flight && units >= min_units;
}

// derive_kernel over `n` entries on `st`.  rec_begin / rec_end: EV_UP1 right before / EV_UP2 right after the launch hcm_stats reports.
int derive_launch(hcm_ctx *c, const UploadEntry *entries, uint32_t n, cudaStream_t st, bool rec_begin, bool rec_end) {
	KernelTune &t = c->tune_derive;
	if (t.in_flight)
		if (int rc = tune_harvest(c))
			return rc;
	if (n && tune_wants_round(c, t, n, c->tune_min_chunks)) {
		// A/B round: baseline, every variant, baseline -- whatever the verdict, the world holds the baseline's bytes
		cudaEvent_t *ev = c->ev_tune[0];
		const uint32_t last = t.n_variants + 1;
		CU(cudaMemsetAsync(c->d_tune_mismatch, 0, HCM_MAX_VARIANTS * sizeof(uint32_t), st));
		for (uint32_t r = 0; r <= last; ++r) {
			const int form = (r == 0 || r == last) ? 0 : (int)r;
			if (r == last && rec_begin)
				CU(cudaEventRecord(c->ev[EV_UP1], st));
			CU(cudaEventRecord(ev[2 * r], st));
			launch_derive(c->w, entries, n, c->d_prop_index, c->d_tables, c->d_rawh, form, true, st);
			CU(cudaEventRecord(ev[2 * r + 1], st));
			if (r == last && rec_end)
				CU(cudaEventRecord(c->ev[EV_UP2], st));
			if (r == 0) {
				launch_derive_digest(c->w, entries, n, c->d_rawh, nullptr, c->d_digest[0], c->d_tune_mismatch, st);
				if (c->tune_fault)
					CU(cudaMemsetAsync(c->d_digest[0], 0xa5, sizeof(unsigned long long), st));
			} else if (r < last)
				launch_derive_digest(c->w, entries, n, c->d_rawh, c->d_digest[0], c->d_digest[1], c->d_tune_mismatch + (r - 1), st);
		}
		CU(cudaGetLastError());
		CU(cudaMemcpyAsync(c->h_tune->mismatch[0], c->d_tune_mismatch, HCM_MAX_VARIANTS * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
		CU(cudaEventRecord(ev[kTuneDone], st));
		t.in_flight = true, t.invalid = false;
		c->stats.kernel_launches += 2 * t.n_variants + 3;
		return HCM_OK;
	}
	if (rec_begin)
		CU(cudaEventRecord(c->ev[EV_UP1], st));
	launch_derive(c->w, entries, n, c->d_prop_index, c->d_tables, c->d_rawh, (int)t.selected, false, st);
	CU(cudaGetLastError());
	if (rec_end)
		CU(cudaEventRecord(c->ev[EV_UP2], st));
	c->stats.kernel_launches += 1;
	return HCM_OK;
}

int launch_batch(hcm_ctx *c) {
	cudaStream_t st = c->stream;
	KernelTune &t = c->tune_mesh;
	if (t.in_flight)
		if (int rc = tune_harvest(c))
			return rc;
	CU(cudaEventRecord(c->ev[EV_B0], st));
	CU(cudaMemsetAsync(c->d_heads, 0, sizeof(BatchHeads), st));
	launch_classify(c->w, c->d_tasks, c->batch_n, c->d_worklist, c->d_heads, c->d_task_out, c->force_all, st);
	c->batch_ab = !c->force_all && !c->d_debug_light && tune_wants_round(c, t, c->batch_n, c->tune_min_tasks);
	if (c->batch_ab) {
		// A/B round over the same work list: baseline, every variant, baseline.  Between two runs the allocation heads and the cursor
		// start over (the work list and its count stay); the last run writes the batch's result
		static_assert(offsetof(BatchHeads, n_work) == 24 && offsetof(BatchHeads, cursor) == 28, "BatchHeads layout");
		cudaEvent_t *ev = c->ev_tune[1];
		const uint32_t last = t.n_variants + 1;
		const unsigned long long *vtx = reinterpret_cast<const unsigned long long *>(c->d_vtx);
		CU(cudaEventRecord(c->ev[EV_BC], st));
		CU(cudaMemsetAsync(c->d_tune_mismatch + HCM_MAX_VARIANTS, 0, HCM_MAX_VARIANTS * sizeof(uint32_t), st));
		for (uint32_t r = 0; r < last; ++r) {
			CU(cudaEventRecord(ev[2 * r], st));
			launch_mesh(c->maps, mesh_params(c, 0), c->grid, (int)r, st);
			CU(cudaEventRecord(ev[2 * r + 1], st));
			if (r == 0) {
				launch_mesh_digest(c->d_task_out, c->d_mesh_out, vtx, c->d_idx, c->batch_n, c->mesh_cap, c->vtx_cap, c->idx_cap, nullptr, c->d_digest[0],
				                   c->d_tune_mismatch, st);
				if (c->tune_fault)
					CU(cudaMemsetAsync(c->d_digest[0], 0xa5, sizeof(unsigned long long), st));
			} else
				launch_mesh_digest(c->d_task_out, c->d_mesh_out, vtx, c->d_idx, c->batch_n, c->mesh_cap, c->vtx_cap, c->idx_cap, c->d_digest[0], c->d_digest[1],
				                   c->d_tune_mismatch + HCM_MAX_VARIANTS + (r - 1), st);
			CU(cudaMemcpyAsync(&c->h_tune->heads[r], c->d_heads, sizeof(BatchHeads), cudaMemcpyDeviceToHost, st));
			CU(cudaMemsetAsync(c->d_heads, 0, offsetof(BatchHeads, n_work), st));
			CU(cudaMemsetAsync(&c->d_heads->cursor, 0, sizeof(unsigned int), st));
		}
		CU(cudaEventRecord(ev[2 * last], st));
	}
	CU(cudaEventRecord(c->ev[EV_B1], st));
	launch_mesh(c->maps, mesh_params(c, 0), c->grid, c->batch_ab ? 0 : (int)t.selected, st);
	CU(cudaGetLastError());
	CU(cudaEventRecord(c->ev[EV_B2], st));
	c->stats.kernel_launches += 2;
	if (c->batch_ab) {
		const uint32_t last = t.n_variants + 1;
		CU(cudaEventRecord(c->ev_tune[1][2 * last + 1], st));
		CU(cudaMemcpyAsync(&c->h_tune->heads[last], c->d_heads, sizeof(BatchHeads), cudaMemcpyDeviceToHost, st));
		CU(cudaMemcpyAsync(c->h_tune->mismatch[1], c->d_tune_mismatch + HCM_MAX_VARIANTS, HCM_MAX_VARIANTS * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
		CU(cudaEventRecord(c->ev_tune[1][kTuneDone], st));
		t.in_flight = true, t.invalid = false;
		c->stats.kernel_launches += 2 * t.n_variants + 2;
	}
	c->ev_batch = true;
	return HCM_OK;
}

// wait for the pending batch; re-run it with larger arenas if an allocation overflowed (never truncate)
int wait_batch(hcm_ctx *c) {
	if (!c->batch_pending && !c->batch_done)
		return fail(c, HCM_ERR_STATE, "no batch submitted");
	if (c->batch_done)
		return HCM_OK;
	for (int attempt = 0; attempt < 4; ++attempt) {
		CU(cudaStreamSynchronize(c->stream));
		CU(cudaMemcpy(c->h_heads, c->d_heads, sizeof(BatchHeads), cudaMemcpyDeviceToHost));
		c->stats.d2h_bytes += sizeof(BatchHeads);
		if (int rc = tune_harvest(c)) // the stream is idle: every round in it has a verdict (an overflowed one is discarded)
			return rc;
		if (!c->h_heads->overflow) {
			c->batch_pending = false, c->batch_done = true;
			return HCM_OK;
		}
		const uint64_t need_v = std::max<uint64_t>(c->h_heads->vtx, c->h_heads->idx / 6 * 4);
		const uint32_t need_m = c->h_heads->meshes;
		int rc = alloc_arenas(c, std::max<uint64_t>(need_v + need_v / 8, c->vtx_cap), std::max<uint32_t>(need_m + need_m / 8 + 64, c->mesh_cap));
		if (rc)
			return rc;
		rc = launch_batch(c);
		if (rc)
			return rc;
	}
	return fail(c, HCM_ERR_CAPACITY, "output arena kept overflowing");
}

} // namespace

extern "C" {
#pragma GCC visibility push(default)

const char *hcm_status_string(int s) {
	switch (s) {
	case HCM_OK: return "ok";
	case HCM_ERR_INVALID: return "invalid argument";
	case HCM_ERR_CUDA: return "CUDA error";
	case HCM_ERR_UNSUPPORTED: return "unsupported registry feature";
	case HCM_ERR_NOT_RESIDENT: return "chunk or neighbour not resident";
	case HCM_ERR_OUT_OF_WINDOW: return "position outside the world window";
	case HCM_ERR_CAPACITY: return "capacity exceeded";
	case HCM_ERR_STATE: return "call sequence error";
	default: return "unknown status";
	}
}

const char *hcm_last_error(const hcm_ctx *c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int hcm_registry_from_blob(const void *blob, size_t size, hcm_registry *out) {
	if (!blob || !out || size < sizeof(hcm_registry_blob_header))
		return HCM_ERR_INVALID;
	const auto *h = (const hcm_registry_blob_header *)blob;
	if (h->magic != HCM_REGISTRY_MAGIC || h->version != HCM_REGISTRY_VERSION)
		return HCM_ERR_INVALID;
	const size_t need = sizeof(*h) + 65536 + (size_t)h->n_props * sizeof(hcm_block_prop) + (size_t)h->n_meshes * sizeof(hcm_custom_mesh) + h->n_textures;
	if (size < need)
		return HCM_ERR_INVALID;
	const uint8_t *p = (const uint8_t *)blob + sizeof(*h);
	out->n_props = h->n_props, out->n_meshes = h->n_meshes, out->n_textures = h->n_textures, out->_pad = 0;
	out->prop_index = p, p += 65536;
	out->props = (const hcm_block_prop *)p, p += (size_t)h->n_props * sizeof(hcm_block_prop);
	out->meshes = (const hcm_custom_mesh *)p, p += (size_t)h->n_meshes * sizeof(hcm_custom_mesh);
	out->tex_flags = p;
	return HCM_OK;
}

void *hcm_host_alloc(size_t bytes) {
	void *p = nullptr;
	if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess)
		return nullptr;
	return p;
}
void hcm_host_free(void *p) {
	if (p)
		cudaFreeHost(p);
}

void hcm_destroy(hcm_ctx *c) {
	if (!c)
		return;
	cudaSetDevice(c->device);
	if (c->stream)
		cudaStreamSynchronize(c->stream);
	void *dev[] = {c->w.prop8, c->w.pass, c->w.torch, c->w.height, c->w.cflags, c->w.slot, c->d_tables, c->d_prop_index, c->d_meshes, c->d_mesh_info, c->d_raw,
	               c->d_rawh, c->d_entries, c->d_all_entries, c->d_tasks, c->d_worklist, c->d_heads, c->d_task_out, c->d_mesh_out, c->d_vtx, c->d_idx, c->d_dyn, c->d_flood_slots, c->d_debug_light, c->d_edits, c->d_edit_columns, c->d_edit_result, c->d_sun_changes, c->d_placements, c->d_digest[0], c->d_digest[1], c->d_tune_mismatch};
	for (void *p : dev)
		if (p)
			cudaFree(p);
	void *host[] = {c->h_heads, c->h_task_out, c->h_mesh_out, c->h_vtx, c->h_idx, c->h_heads_ring, c->h_task_stage, c->h_task_ring[0], c->h_task_ring[1], c->h_tune};
	for (auto &e : c->task_ring_ev)
		if (e)
			cudaEventDestroy(e);
	for (void *p : host)
		if (p)
			cudaFreeHost(p);
	for (auto &e : c->ev)
		if (e)
			cudaEventDestroy(e);
	for (auto &row : c->ev_tune)
		for (auto &e : row)
			if (e)
				cudaEventDestroy(e);
	if (c->ev_peer)
		cudaEventDestroy(c->ev_peer);
	for (auto &e : c->block_events)
		cudaEventDestroy(e);
	for (auto &e : c->sub_events)
		cudaEventDestroy(e);
	if (c->fetch_stream)
		cudaStreamDestroy(c->fetch_stream);
	if (c->copy_stream)
		cudaStreamDestroy(c->copy_stream);
	if (c->stream)
		cudaStreamDestroy(c->stream);
	delete c;
}

int hcm_create(const hcm_registry *registry, const hcm_config *cfg, hcm_ctx **out) {
	if (!out)
		return HCM_ERR_INVALID;
	*out = nullptr;
	hcm_ctx *c = nullptr; // errors before the ctx exists go to the thread-local slot
	if (!cfg || !registry)
		return fail(c, HCM_ERR_INVALID, "null registry/config");
	if (cfg->extent[0] < 3 || cfg->extent[1] < 3 || cfg->extent[2] < 3)
		return fail(c, HCM_ERR_INVALID, "window extent must be at least 3 chunks per axis (a task needs its 26 neighbours)");
	if ((uint64_t)cfg->extent[0] * 32 > 65535 || (uint64_t)cfg->extent[2] * 32 > 65535 || (uint64_t)cfg->extent[1] * 32 > 65535)
		return fail(c, HCM_ERR_INVALID, "window extent too large");

	int n_dev = 0;
	cudaError_t ce = cudaGetDeviceCount(&n_dev);
	if (ce != cudaSuccess || cfg->device < 0 || cfg->device >= n_dev)
		return fail(c, HCM_ERR_CUDA, "no usable CUDA device %d (%s); this library has no CPU fallback", cfg->device,
		            ce != cudaSuccess ? cudaGetErrorString(ce) : "ordinal out of range");
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess || prop.major != 10)
		return fail(c, HCM_ERR_CUDA, "device %d is sm_%d%d; libhcm.so carries sm_100a code only", cfg->device, prop.major, prop.minor);

	hcm_ctx *ctx = new hcm_ctx();
	c = ctx;
	auto bail = [&](int rc) {
		g_create_error = c->err;
		hcm_destroy(c);
		return rc;
	};
	c->device = cfg->device;
	if (const char *ht = getenv("HCM_HOST_THREADS"))
		c->host_threads = (uint32_t)std::max(0, atoi(ht));
	if (const char *ub = getenv("HCM_UPLOAD_BLOCK"))
		c->upload_block = std::max(32u, (uint32_t)atoi(ub) / 32u * 32u); // a multiple of the scan grain
	if (const char *ms = getenv("HCM_MIN_SUB_BATCH"))
		c->min_sub_batch = std::max(1u, (uint32_t)atoi(ms));
	if (const char *fb = getenv("HCM_UPLOAD_FIRST_BLOCK"))
		c->first_block = std::max(32u, (uint32_t)atoi(fb) / 32u * 32u);
	c->first_block = std::min(c->first_block, c->upload_block);
	if (const char *ad = getenv("HCM_UPLOAD_ADAPTIVE"))
		c->adaptive_upload = atoi(ad) != 0;
	if (const char *tu = getenv("HCM_TUNE"))
		c->tune = atoi(tu) != 0;
	c->tune_derive.n_variants = (uint32_t)std::min(derive_variant_count(), HCM_MAX_VARIANTS);
	c->tune_mesh.n_variants = (uint32_t)std::min(mesh_variant_count(), HCM_MAX_VARIANTS);
	if (const char *dv = getenv("HCM_DERIVE_VARIANT"))
		c->tune_derive.forced = std::min(std::max(atoi(dv), 0), (int)c->tune_derive.n_variants), c->tune_derive.selected = (uint32_t)c->tune_derive.forced;
	if (const char *mv = getenv("HCM_MESH_VARIANT"))
		c->tune_mesh.forced = std::min(std::max(atoi(mv), 0), (int)c->tune_mesh.n_variants), c->tune_mesh.selected = (uint32_t)c->tune_mesh.forced;
	if (const char *tc = getenv("HCM_TUNE_MIN_CHUNKS"))
		c->tune_min_chunks = (uint32_t)std::max(1, atoi(tc));
	if (const char *tt = getenv("HCM_TUNE_MIN_TASKS"))
		c->tune_min_tasks = (uint32_t)std::max(1, atoi(tt));
	DevTables tables;
	std::vector<DevMeshInfo> minfo;
	if (int rc = build_tables(c, registry, tables, minfo))
		return bail(rc);

#define CUB(call)                                                                                                      \
	do {                                                                                                               \
		cudaError_t e_ = (call);                                                                                       \
		if (e_ != cudaSuccess) {                                                                                       \
			fail(c, HCM_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);        \
			return bail(HCM_ERR_CUDA);                                                                                 \
		}                                                                                                              \
	} while (0)

	CUB(cudaSetDevice(c->device));
	CUB(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
	CUB(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
	CUB(cudaStreamCreateWithFlags(&c->fetch_stream, cudaStreamNonBlocking));
	for (int i = 0; i < EV_COUNT; ++i)
		CUB(cudaEventCreate(&c->ev[i]));
	CUB(cudaEventCreateWithFlags(&c->ev_peer, cudaEventDisableTiming));

	DevWorld &w = c->w;
	w.ox = cfg->origin[0], w.oy = cfg->origin[1], w.oz = cfg->origin[2];
	w.wx = cfg->extent[0], w.wy = cfg->extent[1], w.wz = cfg->extent[2];
	w.X = 32 * w.wx, w.Z = 32 * w.wz, w.XP = 32 * w.wx + 16;
	c->n_cells = (size_t)w.wx * w.wy * w.wz;
	c->max_chunks = cfg->max_chunks ? cfg->max_chunks : (uint32_t)c->n_cells;
	c->max_tasks = cfg->max_tasks ? cfg->max_tasks : (uint32_t)c->n_cells;
	const size_t n_cols = (size_t)w.wy * w.Z * w.X;
	if (n_cols >= (1ull << 32)) {
		fail(c, HCM_ERR_INVALID, "window too large: %zu columns (the kernels index columns with 32 bits)", n_cols);
		return bail(HCM_ERR_INVALID);
	}

	const size_t n_vox = (size_t)w.wy * 32 * w.Z * w.XP;
	CUB(dev_alloc(c, &w.prop8, n_vox));
	CUB(dev_alloc(c, &w.pass, n_cols));
	if (w.n_levels)
		CUB(dev_alloc(c, &w.torch, n_cols * w.n_levels));
	CUB(dev_alloc(c, &w.height, n_cols));
	CUB(dev_alloc(c, &w.cflags, c->n_cells));
	CUB(dev_alloc(c, &w.slot, c->n_cells));
	CUB(cudaMemsetAsync(w.prop8, 0, n_vox, c->stream));
	CUB(cudaMemsetAsync(w.pass, 0, n_cols * 4, c->stream));
	if (w.n_levels)
		CUB(cudaMemsetAsync(w.torch, 0, n_cols * 4 * w.n_levels, c->stream));
	CUB(cudaMemsetAsync(w.height, 0, n_cols, c->stream));
	CUB(cudaMemsetAsync(w.cflags, 0, c->n_cells * 4, c->stream));
	CUB(cudaMemsetAsync(w.slot, 0xff, c->n_cells * 4, c->stream));
	CUB(dev_alloc(c, &c->d_raw, (size_t)c->max_chunks * HCM_CHUNK_VOXELS));
	CUB(dev_alloc(c, &c->d_rawh, (size_t)c->max_chunks * HCM_CHUNK_COLUMNS));
	CUB(dev_alloc(c, &c->d_entries, (size_t)c->max_chunks));
	w.raw = c->d_raw;

	CUB(dev_alloc(c, &c->d_tables, 1));
	CUB(cudaMemcpyAsync(c->d_tables, &tables, sizeof(tables), cudaMemcpyHostToDevice, c->stream));
	// prop_index[65536] followed by the derive kernel's per-id shortcut: u32[256] indexed by the block id (low byte of a
	// Block): property row | PF flags << 8 | light level << 16 when every meta value maps to the same row, else bit 31
	uint32_t idinfo[256];
	for (uint32_t id = 0; id < 256; ++id) {
		const uint8_t p0 = registry->prop_index[id];
		bool varies = false;
		for (uint32_t meta = 1; meta < 256 && !varies; ++meta)
			varies = registry->prop_index[id | (meta << 8)] != p0;
		idinfo[id] = varies ? 0x80000000u : (uint32_t)p0 | ((uint32_t)tables.pflags[p0] << 8) | ((uint32_t)(tables.pmisc[p0] >> 4) << 16);
	}
	CUB(dev_alloc(c, &c->d_prop_index, 65536 + sizeof(idinfo)));
	CUB(cudaMemcpyAsync(c->d_prop_index, registry->prop_index, 65536, cudaMemcpyHostToDevice, c->stream));
	CUB(cudaMemcpyAsync(c->d_prop_index + 65536, idinfo, sizeof(idinfo), cudaMemcpyHostToDevice, c->stream));
	CUB(dev_alloc(c, &c->d_meshes, std::max<size_t>(registry->n_meshes, 1)));
	CUB(dev_alloc(c, &c->d_mesh_info, std::max<size_t>(registry->n_meshes, 1)));
	if (registry->n_meshes) {
		CUB(cudaMemcpyAsync(c->d_meshes, registry->meshes, sizeof(hcm_custom_mesh) * registry->n_meshes, cudaMemcpyHostToDevice, c->stream));
		CUB(cudaMemcpyAsync(c->d_mesh_info, minfo.data(), sizeof(DevMeshInfo) * registry->n_meshes, cudaMemcpyHostToDevice, c->stream));
	}
	CUB(cudaStreamSynchronize(c->stream)); // `tables`/`minfo` are stack/heap temporaries

	c->slot_of.assign(c->n_cells, -1);
	c->nbr_count.assign(c->n_cells, 0);
	c->free_slots.resize(c->max_chunks);
	for (uint32_t i = 0; i < c->max_chunks; ++i)
		c->free_slots[i] = (int32_t)(c->max_chunks - 1 - i); // pop_back hands out 0,1,2,... -> contiguous uploads

	CUB(dev_alloc(c, &c->d_tasks, c->max_tasks));
	CUB(dev_alloc(c, &c->d_worklist, c->max_tasks));
	CUB(dev_alloc(c, &c->d_heads, 1));
	CUB(dev_alloc(c, &c->d_task_out, c->max_tasks));
	CUB(cudaHostAlloc((void **)&c->h_heads, sizeof(BatchHeads), cudaHostAllocDefault));
	CUB(cudaHostAlloc((void **)&c->h_tune, sizeof(TuneHost), cudaHostAllocDefault));
	memset(c->h_tune, 0, sizeof(TuneHost));
	for (int i = 0; i < 2; ++i)
		CUB(dev_alloc(c, &c->d_digest[i], (size_t)std::max(c->max_chunks, c->max_tasks)));
	CUB(dev_alloc(c, &c->d_tune_mismatch, 2 * HCM_MAX_VARIANTS));
	for (auto &row : c->ev_tune)
		for (auto &e : row)
			CUB(cudaEventCreate(&e));
	for (int i = 0; i < 2; ++i) {
		CUB(cudaHostAlloc((void **)&c->h_task_ring[i], (size_t)c->max_tasks * sizeof(TaskIn), cudaHostAllocDefault));
		CUB(cudaEventCreateWithFlags(&c->task_ring_ev[i], cudaEventDisableTiming));
	}
	const uint64_t v0 = cfg->vertex_capacity ? cfg->vertex_capacity : std::max<uint64_t>(4ull << 20, (uint64_t)c->max_tasks * 512);
	if (int rc = alloc_arenas(c, v0, 2 * c->max_tasks + 64))
		return bail(rc);

	CUB(mesh_kernel_prepare());
	c->grid = prop.multiProcessorCount; // one persistent CTA per SM (shared-memory bound)
	CUB(dev_alloc(c, &c->d_dyn, (size_t)c->grid * kLightBytes));
	CUB(dev_alloc(c, &c->d_flood_slots, (size_t)kThreads));
	launch_flood_slots(c->d_flood_slots, c->stream);
	CUB(cudaGetLastError());

	// TMA descriptor over the dense property grid: dims (x, z, y), box 48 x 36 x 34, out-of-window reads give 0 (= air)
	{
		void *fn = nullptr;
		cudaDriverEntryPointQueryResult qres;
		CUB(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
		if (!fn || qres != cudaDriverEntryPointSuccess) {
			fail(c, HCM_ERR_CUDA, "cuTensorMapEncodeTiled not available in this driver");
			return bail(HCM_ERR_CUDA);
		}
		const cuuint64_t dims[3] = {(cuuint64_t)w.XP, (cuuint64_t)w.Z, (cuuint64_t)w.wy * 32};
		const cuuint64_t strides[2] = {(cuuint64_t)w.XP, (cuuint64_t)w.XP * w.Z};
		const cuuint32_t box[3] = {(cuuint32_t)kVX, (cuuint32_t)kVZ, (cuuint32_t)kVY};
		const cuuint32_t estr[3] = {1, 1, 1};
		CUresult r = ((PFN_encodeTiled)fn)(&c->maps.vox, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, w.prop8, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
		                                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
		// flood inputs: the 62 x 62 columns of a task's light domain over three chunk layers, widened to 64 columns so that the box
		// starts on a 16-byte boundary (x = 32 cx - 16); dims (x, z, chunk layer)
		const cuuint64_t fdims[3] = {(cuuint64_t)w.X, (cuuint64_t)w.Z, (cuuint64_t)w.wy};
		const cuuint32_t fbox[3] = {(cuuint32_t)kFloodBoxX, (cuuint32_t)kFD, 3u};
		const cuuint64_t pstr[2] = {(cuuint64_t)w.X * 4, (cuuint64_t)w.X * w.Z * 4};
		const cuuint64_t hstr[2] = {(cuuint64_t)w.X, (cuuint64_t)w.X * w.Z};
		if (r == CUDA_SUCCESS)
			r = ((PFN_encodeTiled)fn)(&c->maps.pass, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, w.pass, fdims, pstr, fbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
			                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
		if (r == CUDA_SUCCESS)
			r = ((PFN_encodeTiled)fn)(&c->maps.height, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, w.height, fdims, hstr, fbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
			                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
		if (r != CUDA_SUCCESS) {
			fail(c, HCM_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
			return bail(HCM_ERR_CUDA);
		}
	}
#undef CUB
	*out = c;
	return HCM_OK;
}

int hcm_is_resident(const hcm_ctx *c, const int16_t pos[3]) {
	int cx, cy, cz;
	if (!c || !pos || !in_window(c, pos, &cx, &cy, &cz))
		return 0;
	return c->slot_of[cell(c, cx, cy, cz)] >= 0 ? 1 : 0;
}

// Hooks of upload_impl: hcm_upload_and_mesh submits mesh sub-batches between the upload blocks.
struct UploadHooks {
	std::function<int(const std::vector<UploadEntry> &, uint32_t n_blocks)> assigned; // residency is updated, nothing enqueued yet
	std::function<int(uint32_t blk, uint32_t n_blocks)> block_enqueued;               // copy + derive of block blk are in the streams
};

static int upload_impl(hcm_ctx *c, uint32_t n, const int16_t *pos, const uint16_t *blocks, const int8_t *heights, const int32_t *hint,
                       const UploadHooks *hooks) {
	if (!pos || !blocks || !heights)
		return fail(c, HCM_ERR_INVALID, "null upload arrays");
	if (n > c->max_chunks)
		return fail(c, HCM_ERR_CAPACITY, "upload of %u chunks exceeds max_chunks %u", n, c->max_chunks);
	CU(cudaSetDevice(c->device));
	// Uniform chunks (all air / all stone: most of a load sphere) cross the bus as one Block value instead of 64 KB.
	// Pipeline per block of upload_block chunks: host threads scan block b+1 (early exit at the first difference) while
	// the copy engine moves block b and the derive kernel runs on block b-1.  The scan starts before anything else: it
	// only reads the caller's arrays, and it is the longest stage of a large upload.
	//
	// The scan and the copy engine balance themselves: a work item (kScanGrain chunks) belongs to whoever claims it first.
	// Scan threads claim items front to back.  The uploading thread, when it reaches a block whose scan is not complete while
	// the copy engine has nothing left to do, claims the block's unclaimed items itself and sends them as they are -- the
	// derive kernel recognises a uniform chunk on the device as well (its AND/OR test), at HBM speed.  With many scan threads
	// per GPU next to nothing goes unscanned; with 8 GPUs sharing one host's cores (3 scan threads each) most of the world does,
	// and the upload runs at bus speed instead of at the speed of the scan.
	// Caller's knowledge (hcm_upload_chunks_hinted) short-cuts both: a chunk hinted uniform is neither read nor copied, a chunk
	// hinted not-uniform is copied unread.
	const uint32_t n_items = (n + kScanGrain - 1) / kScanGrain;
	enum : uint8_t { IT_FREE = 0, IT_BUSY = 1, IT_DONE = 2 };
	std::vector<std::atomic<uint8_t>> item_state(n_items);
	for (auto &d : item_state)
		d.store(IT_FREE, std::memory_order_relaxed);
	std::vector<int32_t> uniform_value(n); // Block value of a uniform chunk, else -1; written by the item's owner before IT_DONE
	std::atomic<uint32_t> next{0};
	std::atomic<bool> abort_scan{false};
	int32_t *uni = uniform_value.data();
	// resolve one item; `scan` = read the chunks that carry no hint
	auto resolve_item = [uni, n, blocks, hint](uint32_t it, bool scan) {
		const uint32_t lo = it * kScanGrain, hi = std::min(n, lo + kScanGrain);
		for (uint32_t i = lo; i < hi; ++i) {
			const int32_t h = hint ? hint[i] : HCM_HINT_UNKNOWN;
			if (h >= 0)
				uni[i] = h & 0xffff;
			else if (h == HCM_HINT_NOT_UNIFORM || !scan)
				uni[i] = -1;
			else {
				const uint16_t *b = blocks + (size_t)i * HCM_CHUNK_VOXELS;
				uni[i] = chunk_is_uniform(b) ? (int32_t)b[0] : -1;
			}
		}
	};
	auto scan_work = [&item_state, &next, &abort_scan, resolve_item, n_items]() {
		for (;;) {
			const uint32_t it = next.fetch_add(1);
			if (it >= n_items || abort_scan.load(std::memory_order_relaxed))
				break;
			uint8_t expect = IT_FREE;
			if (!item_state[it].compare_exchange_strong(expect, IT_BUSY, std::memory_order_acq_rel))
				continue; // the uploading thread took it unscanned
			resolve_item(it, true);
			item_state[it].store(IT_DONE, std::memory_order_release);
		}
	};
	const unsigned hw = std::max(2u, std::thread::hardware_concurrency());
	const uint32_t want_threads = c->host_threads ? c->host_threads : std::min<uint32_t>(hw - 1, 15);
	const bool threaded = n >= 256 && want_threads > 0;
	if (threaded) {
		if (!c->pool || c->pool->threads.size() != want_threads) {
			c->pool.reset(new WorkerPool);
			c->pool->start(want_threads);
		}
		c->pool->launch(scan_work);
	}
	struct ScanGuard { // the workers reference this frame: never leave it while they run
		hcm_ctx *c;
		bool active;
		std::atomic<bool> &abort_scan;
		~ScanGuard() {
			if (active) {
				abort_scan.store(true);
				c->pool->wait();
			}
		}
	} scan_guard{c, threaded, abort_scan};
	const bool adaptive = threaded && c->adaptive_upload;

	std::vector<UploadEntry> entries(n);
	std::vector<int32_t> fresh; // slots taken in this call (rolled back on error)
	auto roll_back = [&]() {
		for (auto it = fresh.rbegin(); it != fresh.rend(); ++it) {
			c->free_slots.push_back(c->slot_of[*it]);
			c->slot_of[*it] = -1;
			const int fx = *it % c->w.wx, fz = (*it / c->w.wx) % c->w.wz, fy = *it / (c->w.wx * c->w.wz);
			bump_neighbours(c, fx, fy, fz, -1);
		}
	};
	for (uint32_t i = 0; i < n; ++i) {
		int cx, cy, cz;
		int rc = HCM_OK;
		if (!in_window(c, pos + 3 * i, &cx, &cy, &cz))
			rc = fail(c, HCM_ERR_OUT_OF_WINDOW, "chunk (%d,%d,%d) is outside the window", pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
		int32_t slot = -1;
		if (!rc) {
			slot = c->slot_of[cell(c, cx, cy, cz)];
			if (slot < 0) {
				if (c->free_slots.empty())
					rc = fail(c, HCM_ERR_CAPACITY, "more than %u resident chunks", c->max_chunks);
				else {
					slot = c->free_slots.back();
					c->free_slots.pop_back();
					c->slot_of[cell(c, cx, cy, cz)] = slot;
					fresh.push_back((int32_t)cell(c, cx, cy, cz));
					bump_neighbours(c, cx, cy, cz, +1);
				}
			}
		}
		if (rc) {
			roll_back();
			return rc;
		}
		entries[i] = UploadEntry{slot, (int16_t)cx, (int16_t)cy, (int16_t)cz, 0, 0};
	}
	// block boundaries are multiples of kScanGrain (upload_block and first_block are)
	const BlockLayout layout{n, adaptive ? std::min(c->first_block, c->upload_block) : c->upload_block, c->upload_block};
	c->layout = layout;
	const uint32_t n_blocks = layout.count();
	if (hooks && hooks->assigned)
		if (int rc = hooks->assigned(entries, n_blocks)) {
			roll_back();
			return rc;
		}
	cudaStream_t st = c->stream, cs = c->copy_stream;
	CU(cudaEventRecord(c->ev[EV_UP0], st));
	CU(cudaStreamWaitEvent(cs, c->ev[EV_UP0], 0)); // earlier kernels may still read the raw slots
	while (c->block_events.size() < n_blocks) {
		cudaEvent_t e;
		CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
		c->block_events.push_back(e);
	}
	UploadEntry *ent = entries.data();
	if (!threaded)
		scan_work();

	uint64_t copied = 0;
	uint32_t n_unscanned = 0;
	std::vector<void *> cp_dst, cp_src;
	std::vector<size_t> cp_size;
	for (uint32_t blk = 0; blk < n_blocks; ++blk) {
		const uint32_t lo = layout.lo(blk), hi = layout.hi(blk);
		bool engine_idle = false; // the copy stream has finished everything enqueued so far (it never becomes busy again before this block)
		for (uint32_t it = lo / kScanGrain; it < (hi + kScanGrain - 1) / kScanGrain; ++it)
			for (;;) {
				const uint8_t s = item_state[it].load(std::memory_order_acquire);
				if (s == IT_DONE)
					break;
				if (s == IT_FREE && adaptive) {
					if (!engine_idle) {
						engine_idle = blk == 0 || cudaEventQuery(c->block_events[blk - 1]) == cudaSuccess;
						if (!engine_idle)
							(void)cudaGetLastError(); // "not ready" is an answer, not an error: it must not surface at the next cudaGetLastError()
					}
					uint8_t expect = IT_FREE;
					if (engine_idle && item_state[it].compare_exchange_strong(expect, IT_BUSY, std::memory_order_acq_rel)) {
						resolve_item(it, false);
						item_state[it].store(IT_DONE, std::memory_order_release);
						n_unscanned += std::min(n, (it + 1) * kScanGrain) - it * kScanGrain;
						break;
					}
				}
				std::this_thread::yield();
			}
		for (uint32_t i = lo; i < hi; ++i)
			if (uni[i] >= 0)
				ent[i].flags |= UE_UNIFORM, ent[i].value = (uint16_t)uni[i];
		// heights for every chunk, blocks for the non-uniform ones; one copy per run of consecutive slots, all runs of the
		// block submitted with ONE cudaMemcpyBatchAsync (thousands of single calls would cost more than the transfer)
		cp_dst.clear(), cp_src.clear(), cp_size.clear();
		for (uint32_t i = lo; i < hi;) {
			uint32_t j = i + 1;
			while (j < hi && ent[j].slot == ent[j - 1].slot + 1)
				++j;
			cp_dst.push_back(c->d_rawh + (size_t)ent[i].slot * HCM_CHUNK_COLUMNS), cp_src.push_back((void *)(heights + (size_t)i * HCM_CHUNK_COLUMNS));
			cp_size.push_back((size_t)(j - i) * HCM_CHUNK_COLUMNS);
			for (uint32_t a = i; a < j;) {
				if (ent[a].flags & UE_UNIFORM) {
					++a;
					continue;
				}
				uint32_t b = a + 1;
				while (b < j && !(ent[b].flags & UE_UNIFORM))
					++b;
				cp_dst.push_back(c->d_raw + (size_t)ent[a].slot * HCM_CHUNK_VOXELS), cp_src.push_back((void *)(blocks + (size_t)a * HCM_CHUNK_VOXELS));
				cp_size.push_back((size_t)(b - a) * HCM_CHUNK_VOXELS * 2);
				a = b;
			}
			i = j;
		}
		for (size_t k = 0; k < cp_size.size(); ++k)
			copied += cp_size[k];
		bool batched = false;
		if (c->use_batch_copy && cp_size.size() > 4) {
			cudaMemcpyAttributes attr{};
			attr.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
			size_t attr_idx = 0, fail_idx = 0;
			cudaError_t be = cudaMemcpyBatchAsync(cp_dst.data(), cp_src.data(), cp_size.data(), cp_size.size(), &attr, &attr_idx, 1, &fail_idx, cs);
			if (be == cudaSuccess)
				batched = true;
			else {
				(void)cudaGetLastError();
				c->use_batch_copy = false; // driver/runtime without batch copies: fall back for good
			}
		}
		if (!batched)
			for (size_t k = 0; k < cp_size.size(); ++k)
				CU(cudaMemcpyAsync(cp_dst[k], cp_src[k], cp_size[k], cudaMemcpyHostToDevice, cs));
		CU(cudaMemcpyAsync(c->d_entries + lo, ent + lo, (size_t)(hi - lo) * sizeof(UploadEntry), cudaMemcpyHostToDevice, cs)); // pageable: staged before return
		CU(cudaEventRecord(c->block_events[blk], cs));
		CU(cudaStreamWaitEvent(st, c->block_events[blk], 0));
		if (int rc = derive_launch(c, c->d_entries + lo, hi - lo, st, blk == 0, blk + 1 == n_blocks))
			return rc;
		if (hooks && hooks->block_enqueued)
			if (int rc = hooks->block_enqueued(blk, n_blocks))
				return rc;
	}
	c->stats.h2d_bytes += copied + (uint64_t)n * sizeof(UploadEntry);
	c->stats.upload_unscanned_chunks = n_unscanned;
	c->ev_upload = true;
	return HCM_OK;
}

int hcm_upload_chunks(hcm_ctx *c, uint32_t n, const int16_t *pos, const uint16_t *blocks, const int8_t *heights) {
	if (!c)
		return HCM_ERR_INVALID;
	if (n == 0)
		return HCM_OK;
	return upload_impl(c, n, pos, blocks, heights, nullptr, nullptr);
}

int hcm_upload_chunks_hinted(hcm_ctx *c, uint32_t n, const int16_t *pos, const uint16_t *blocks, const int8_t *heights, const int32_t *uniform_hint) {
	if (!c)
		return HCM_ERR_INVALID;
	if (n == 0)
		return HCM_OK;
	if (uniform_hint)
		for (uint32_t i = 0; i < n; ++i)
			if (uniform_hint[i] > 0xffff || uniform_hint[i] < HCM_HINT_NOT_UNIFORM)
				return fail(c, HCM_ERR_INVALID, "uniform_hint[%u] = %d is neither a Block value nor HCM_HINT_*", i, uniform_hint[i]);
	return upload_impl(c, n, pos, blocks, heights, uniform_hint, nullptr);
}

int hcm_copy_chunks_from(hcm_ctx *c, hcm_ctx *src, uint32_t n, const int16_t *pos) {
	if (!c || !src)
		return HCM_ERR_INVALID;
	if (n == 0)
		return HCM_OK;
	if (!pos)
		return fail(c, HCM_ERR_INVALID, "null positions");
	if (c == src)
		return fail(c, HCM_ERR_INVALID, "source and destination are the same context");
	if (n > c->max_chunks)
		return fail(c, HCM_ERR_CAPACITY, "copy of %u chunks exceeds max_chunks %u", n, c->max_chunks);
	std::vector<UploadEntry> entries(n);
	std::vector<int32_t> src_slot(n), fresh;
	auto roll_back = [&]() {
		for (auto it = fresh.rbegin(); it != fresh.rend(); ++it) {
			c->free_slots.push_back(c->slot_of[*it]);
			c->slot_of[*it] = -1;
			const int fx = *it % c->w.wx, fz = (*it / c->w.wx) % c->w.wz, fy = *it / (c->w.wx * c->w.wz);
			bump_neighbours(c, fx, fy, fz, -1);
		}
	};
	for (uint32_t i = 0; i < n; ++i) {
		int sx, sy, sz, cx, cy, cz;
		int rc = HCM_OK;
		if (!in_window(src, pos + 3 * i, &sx, &sy, &sz) || src->slot_of[cell(src, sx, sy, sz)] < 0)
			rc = fail(c, HCM_ERR_NOT_RESIDENT, "chunk (%d,%d,%d) is not resident in the source context", pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
		else if (!in_window(c, pos + 3 * i, &cx, &cy, &cz))
			rc = fail(c, HCM_ERR_OUT_OF_WINDOW, "chunk (%d,%d,%d) is outside the window", pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
		int32_t slot = -1;
		if (!rc) {
			src_slot[i] = src->slot_of[cell(src, sx, sy, sz)];
			slot = c->slot_of[cell(c, cx, cy, cz)];
			if (slot < 0) {
				if (c->free_slots.empty())
					rc = fail(c, HCM_ERR_CAPACITY, "more than %u resident chunks", c->max_chunks);
				else {
					slot = c->free_slots.back();
					c->free_slots.pop_back();
					c->slot_of[cell(c, cx, cy, cz)] = slot;
					fresh.push_back((int32_t)cell(c, cx, cy, cz));
					bump_neighbours(c, cx, cy, cz, +1);
				}
			}
		}
		if (rc) {
			roll_back();
			return rc;
		}
		entries[i] = UploadEntry{slot, (int16_t)cx, (int16_t)cy, (int16_t)cz, 0, 0};
	}
	// direct access between the two GPUs (NVLink / NVSwitch) where the topology offers it; without it the driver stages the copy
	if (src->device != c->device && src->device < 64 && !((c->peers_tried >> src->device) & 1)) {
		c->peers_tried |= 1ull << src->device;
		int can = 0;
		CU(cudaSetDevice(c->device));
		if (cudaDeviceCanAccessPeer(&can, c->device, src->device) == cudaSuccess && can)
			if (cudaDeviceEnablePeerAccess(src->device, 0) != cudaSuccess)
				(void)cudaGetLastError(); // already enabled by somebody else in this process: fine
	}
	// src's stream -> event -> dst's stream: the copies read what src had enqueued (edits, derives) before this call
	CU(cudaSetDevice(src->device));
	CU(cudaEventRecord(src->ev_peer, src->stream));
	CU(cudaSetDevice(c->device));
	cudaStream_t st = c->stream;
	CU(cudaStreamWaitEvent(st, src->ev_peer, 0));
	for (uint32_t i = 0; i < n;) { // one copy per run of chunks whose slots are consecutive on both sides
		uint32_t j = i + 1;
		while (j < n && entries[j].slot == entries[j - 1].slot + 1 && src_slot[j] == src_slot[j - 1] + 1)
			++j;
		CU(cudaMemcpyPeerAsync(c->d_raw + (size_t)entries[i].slot * HCM_CHUNK_VOXELS, c->device, src->d_raw + (size_t)src_slot[i] * HCM_CHUNK_VOXELS, src->device,
		                       (size_t)(j - i) * HCM_CHUNK_VOXELS * 2, st));
		CU(cudaMemcpyPeerAsync(c->d_rawh + (size_t)entries[i].slot * HCM_CHUNK_COLUMNS, c->device, src->d_rawh + (size_t)src_slot[i] * HCM_CHUNK_COLUMNS,
		                       src->device, (size_t)(j - i) * HCM_CHUNK_COLUMNS, st));
		i = j;
	}
	CU(cudaEventRecord(c->ev_peer, st)); // src must not overwrite the slots (a later edit) before they have been read
	CU(cudaMemcpyAsync(c->d_entries, entries.data(), (size_t)n * sizeof(UploadEntry), cudaMemcpyHostToDevice, st)); // pageable: staged before return
	if (int rc = derive_launch(c, c->d_entries, n, st, false, false))
		return rc;
	c->stats.p2p_bytes += (uint64_t)n * (HCM_CHUNK_VOXELS * 2 + HCM_CHUNK_COLUMNS);
	c->stats.h2d_bytes += (uint64_t)n * sizeof(UploadEntry);
	CU(cudaSetDevice(src->device));
	CU(cudaStreamWaitEvent(src->stream, c->ev_peer, 0));
	CU(cudaSetDevice(c->device));
	return HCM_OK;
}

int hcm_remove_chunks(hcm_ctx *c, uint32_t n, const int16_t *pos) {
	if (!c)
		return HCM_ERR_INVALID;
	if (n == 0)
		return HCM_OK;
	if (!pos)
		return fail(c, HCM_ERR_INVALID, "null positions");
	CU(cudaSetDevice(c->device));
	std::vector<UploadEntry> entries;
	for (uint32_t i = 0; i < n; ++i) {
		int cx, cy, cz;
		if (!in_window(c, pos + 3 * i, &cx, &cy, &cz))
			continue;
		int32_t &slot = c->slot_of[cell(c, cx, cy, cz)];
		if (slot < 0)
			continue;
		entries.push_back(UploadEntry{slot, (int16_t)cx, (int16_t)cy, (int16_t)cz, 0, 0});
		c->free_slots.push_back(slot);
		slot = -1;
		bump_neighbours(c, cx, cy, cz, -1);
	}
	if (entries.empty())
		return HCM_OK;
	CU(cudaMemcpyAsync(c->d_entries, entries.data(), entries.size() * sizeof(UploadEntry), cudaMemcpyHostToDevice, c->stream));
	launch_clear(c->w, c->d_entries, (uint32_t)entries.size(), c->stream);
	CU(cudaGetLastError());
	c->stats.kernel_launches += 1;
	return HCM_OK;
}

int hcm_rederive_chunks(hcm_ctx *c, uint32_t n, const int16_t *pos) {
	if (!c)
		return HCM_ERR_INVALID;
	if (n && !pos)
		return fail(c, HCM_ERR_INVALID, "null positions");
	CU(cudaSetDevice(c->device));
	std::vector<UploadEntry> entries;
	if (n == 0) {
		cudaStream_t st0 = c->stream;
		if (c->all_entries_version != c->residency_version) { // the list only changes with residency: keep it on the device
			for (int cy = 0; cy < c->w.wy; ++cy)
				for (int cz = 0; cz < c->w.wz; ++cz)
					for (int cx = 0; cx < c->w.wx; ++cx) {
						const int32_t slot = c->slot_of[cell(c, cx, cy, cz)];
						if (slot >= 0)
							entries.push_back(UploadEntry{slot, (int16_t)cx, (int16_t)cy, (int16_t)cz, 0, 0});
					}
			if (!c->d_all_entries)
				CU(dev_alloc(c, &c->d_all_entries, (size_t)c->max_chunks));
			if (!entries.empty())
				CU(cudaMemcpyAsync(c->d_all_entries, entries.data(), entries.size() * sizeof(UploadEntry), cudaMemcpyHostToDevice, st0));
			c->stats.h2d_bytes += entries.size() * sizeof(UploadEntry);
			c->n_all_entries = (uint32_t)entries.size();
			c->all_entries_version = c->residency_version;
		}
		if (!c->n_all_entries)
			return HCM_OK;
		CU(cudaEventRecord(c->ev[EV_UP0], st0));
		if (int rc = derive_launch(c, c->d_all_entries, c->n_all_entries, st0, true, true))
			return rc;
		c->ev_upload = true;
		return HCM_OK;
	} else {
		for (uint32_t i = 0; i < n; ++i) {
			int cx, cy, cz;
			if (!in_window(c, pos + 3 * i, &cx, &cy, &cz))
				return fail(c, HCM_ERR_OUT_OF_WINDOW, "chunk (%d,%d,%d) is outside the window", pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
			const int32_t slot = c->slot_of[cell(c, cx, cy, cz)];
			if (slot < 0)
				return fail(c, HCM_ERR_NOT_RESIDENT, "chunk (%d,%d,%d) is not resident", pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
			entries.push_back(UploadEntry{slot, (int16_t)cx, (int16_t)cy, (int16_t)cz, 0, 0});
		}
	}
	if (entries.empty())
		return HCM_OK;
	cudaStream_t st = c->stream;
	CU(cudaEventRecord(c->ev[EV_UP0], st));
	CU(cudaMemcpyAsync(c->d_entries, entries.data(), entries.size() * sizeof(UploadEntry), cudaMemcpyHostToDevice, st));
	c->stats.h2d_bytes += entries.size() * sizeof(UploadEntry);
	if (int rc = derive_launch(c, c->d_entries, (uint32_t)entries.size(), st, true, true))
		return rc;
	c->ev_upload = true;
	return HCM_OK;
}

// ChunkShouldRemesh (client/include/client/ChunkMeshTask.inl:9-15): a block at `rel` (relative to a chunk's origin) can
// reach that chunk's mesh iff its L1 distance to the box [-1,33]^3 is below 15 (IsBorderLightInterfere at level 15)
static bool chunk_should_remesh(const int rel[3]) {
	int dist = 0;
	for (int a = 0; a < 3; ++a) {
		if (rel[a] < -1)
			dist += -1 - rel[a];
		else if (rel[a] >= HCM_CHUNK_SIZE + 1)
			dist += rel[a] - (HCM_CHUNK_SIZE + 1);
	}
	return 15 > dist;
}

int hcm_set_blocks(hcm_ctx *c, uint32_t n, const hcm_block_edit *edits, hcm_edit_result *out) {
	if (!c || !out)
		return HCM_ERR_INVALID;
	memset(out, 0, sizeof(*out));
	if (n == 0)
		return HCM_OK;
	if (!edits)
		return fail(c, HCM_ERR_INVALID, "null edits");
	CU(cudaSetDevice(c->device));
	for (uint32_t i = 0; i < n; ++i) {
		int cx, cy, cz;
		if (!in_window(c, edits[i].chunk, &cx, &cy, &cz))
			return fail(c, HCM_ERR_OUT_OF_WINDOW, "edit %u: chunk (%d,%d,%d) is outside the window", i, edits[i].chunk[0], edits[i].chunk[1], edits[i].chunk[2]);
		if (c->slot_of[cell(c, cx, cy, cz)] < 0)
			return fail(c, HCM_ERR_NOT_RESIDENT, "edit %u: chunk (%d,%d,%d) is not resident", i, edits[i].chunk[0], edits[i].chunk[1], edits[i].chunk[2]);
		if (edits[i].index >= HCM_CHUNK_VOXELS)
			return fail(c, HCM_ERR_INVALID, "edit %u: block index %u out of range", i, edits[i].index);
	}
	// order: world column (chunk x, chunk z, x, z), highest chunk first; one entry per voxel, the last one given wins
	// (the reference keeps one pending update per block index, ChunkTaskPool's set-block map)
	std::vector<uint32_t> order(n);
	for (uint32_t i = 0; i < n; ++i)
		order[i] = i;
	auto key_of = [](const hcm_block_edit &e) {
		return ((uint64_t)(uint16_t)e.chunk[0] << 32) | ((uint64_t)(uint16_t)e.chunk[2] << 16) | (uint64_t)((((e.index >> 5) & 31u) << 5) | (e.index & 31u));
	};
	std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
		const uint64_t ka = key_of(edits[a]), kb = key_of(edits[b]);
		if (ka != kb)
			return ka < kb;
		if (edits[a].chunk[1] != edits[b].chunk[1])
			return edits[a].chunk[1] > edits[b].chunk[1];
		if (edits[a].index != edits[b].index)
			return edits[a].index > edits[b].index;
		return a < b;
	});
	std::vector<hcm_block_edit> sorted;
	std::vector<uint2> columns; // per world column: first edit, number of edits
	sorted.reserve(n);
	for (uint32_t k = 0; k < n; ++k) {
		const hcm_block_edit &e = edits[order[k]];
		if (k + 1 < n) { // the same voxel again, later in the caller's order: that one wins
			const hcm_block_edit &f = edits[order[k + 1]];
			if (key_of(e) == key_of(f) && e.chunk[1] == f.chunk[1] && e.index == f.index)
				continue;
		}
		if (sorted.empty() || key_of(sorted.back()) != key_of(e))
			columns.push_back(make_uint2((unsigned)sorted.size(), 0u));
		sorted.push_back(e);
		columns.back().y++;
	}
	const uint32_t m = (uint32_t)sorted.size(), ncol = (uint32_t)columns.size();
	if (m > c->edit_cap) {
		for (void *p : {(void *)c->d_edits, (void *)c->d_edit_columns, (void *)c->d_edit_result})
			if (p)
				cudaFree(p);
		c->d_edits = nullptr, c->d_edit_columns = nullptr, c->d_edit_result = nullptr;
		const uint32_t cap = std::max(1024u, m + m / 2);
		CU(cudaMalloc((void **)&c->d_edits, (size_t)cap * sizeof(hcm_block_edit)));
		CU(cudaMalloc((void **)&c->d_edit_columns, (size_t)cap * sizeof(uint2)));
		CU(cudaMalloc((void **)&c->d_edit_result, ((size_t)cap + 1) * sizeof(uint32_t)));
		c->edit_cap = cap;
	}
	const uint32_t sun_need = m * (uint32_t)c->w.wy; // a chain visits each chunk of a column at most once per edit
	if (sun_need > c->sun_cap) {
		if (c->d_sun_changes)
			cudaFree(c->d_sun_changes);
		c->d_sun_changes = nullptr;
		CU(cudaMalloc((void **)&c->d_sun_changes, (size_t)sun_need * sizeof(hcm_sunlight_change)));
		c->sun_cap = sun_need;
	}
	cudaStream_t st = c->stream;
	uint32_t *d_count = c->d_edit_result + c->edit_cap;
	CU(cudaMemcpyAsync(c->d_edits, sorted.data(), (size_t)m * sizeof(hcm_block_edit), cudaMemcpyHostToDevice, st));
	CU(cudaMemcpyAsync(c->d_edit_columns, columns.data(), (size_t)ncol * sizeof(uint2), cudaMemcpyHostToDevice, st));
	CU(cudaMemsetAsync(d_count, 0, sizeof(uint32_t), st));
	launch_edit(c->w, c->d_edits, m, c->d_edit_columns, ncol, c->d_edit_result, c->d_prop_index, c->d_tables, c->d_rawh, c->d_sun_changes, d_count, c->sun_cap, st);
	CU(cudaGetLastError());
	c->stats.kernel_launches += 2;
	std::vector<uint32_t> result(m);
	uint32_t n_changes = 0;
	CU(cudaMemcpyAsync(result.data(), c->d_edit_result, (size_t)m * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
	CU(cudaMemcpyAsync(&n_changes, d_count, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	if (n_changes > c->sun_cap)
		return fail(c, HCM_ERR_STATE, "internal: sunlight change list overflow");
	c->edit_sunlight.resize(n_changes);
	if (n_changes)
		CU(cudaMemcpy(c->edit_sunlight.data(), c->d_sun_changes, (size_t)n_changes * sizeof(hcm_sunlight_change), cudaMemcpyDeviceToHost));
	c->stats.h2d_bytes += (uint64_t)m * sizeof(hcm_block_edit) + (uint64_t)ncol * sizeof(uint2);
	c->stats.d2h_bytes += (uint64_t)m * 4 + 4 + (uint64_t)n_changes * sizeof(hcm_sunlight_change);

	// fan-out on the host: which chunks to derive again, which mesh tasks the reference pushes
	std::vector<uint8_t> mark(c->n_cells, 0); // bit0: derive again, bit1: remesh pushed by a block edit, bit2: by a sunlight change
	auto mark_remesh = [&](int x, int y, int z, uint8_t why) {
		if (x >= 0 && y >= 0 && z >= 0 && x < c->w.wx && y < c->w.wy && z < c->w.wz)
			mark[cell(c, x, y, z)] |= why;
		else
			out->n_pending++;
	};
	for (uint32_t k = 0; k < m; ++k) {
		if (!((result[k] >> 16) & 1u))
			continue;
		out->n_changed_blocks++;
		const hcm_block_edit &e = sorted[k];
		const int cx = e.chunk[0] - c->w.ox, cy = e.chunk[1] - c->w.oy, cz = e.chunk[2] - c->w.oz;
		const int bp[3] = {(int)(e.index & 31u), (int)(e.index >> 10), (int)((e.index >> 5) & 31u)};
		mark[cell(c, cx, cy, cz)] |= 3;
		for (int dy = -1; dy <= 1; ++dy)
			for (int dz = -1; dz <= 1; ++dz)
				for (int dx = -1; dx <= 1; ++dx) {
					if (!dx && !dy && !dz)
						continue;
					const int rel[3] = {bp[0] - dx * HCM_CHUNK_SIZE, bp[1] - dy * HCM_CHUNK_SIZE, bp[2] - dz * HCM_CHUNK_SIZE};
					if (chunk_should_remesh(rel)) // ChunkSetBlockTask.cpp:100-107
						mark_remesh(cx + dx, cy + dy, cz + dz, 2);
				}
	}
	for (const hcm_sunlight_change &s : c->edit_sunlight) {
		const int cx = s.chunk[0] - c->w.ox, cy = s.chunk[1] - c->w.oy, cz = s.chunk[2] - c->w.oz;
		mark[cell(c, cx, cy, cz)] |= 1;
		for (int dy = -1; dy <= 1; ++dy) // ChunkSetSunlightTask.cpp:66-78: the test is against the chunk itself for every neighbour
			for (int dz = -1; dz <= 1; ++dz)
				for (int dx = -1; dx <= 1; ++dx)
					mark_remesh(cx + dx, cy + dy, cz + dz, 4);
	}
	std::vector<int16_t> derive_pos;
	c->edit_remesh.clear(), c->edit_remesh_flags.clear();
	for (int y = 0; y < c->w.wy; ++y)
		for (int z = 0; z < c->w.wz; ++z)
			for (int x = 0; x < c->w.wx; ++x) {
				const uint8_t f = mark[cell(c, x, y, z)];
				if (!f)
					continue;
				const int16_t p[3] = {(int16_t)(x + c->w.ox), (int16_t)(y + c->w.oy), (int16_t)(z + c->w.oz)};
				if (f & 1)
					derive_pos.insert(derive_pos.end(), p, p + 3);
				if (f & 6) {
					const bool interior = x >= 1 && y >= 1 && z >= 1 && x <= c->w.wx - 2 && y <= c->w.wy - 2 && z <= c->w.wz - 2;
					if (interior && c->nbr_count[cell(c, x, y, z)] == 27) {
						c->edit_remesh.insert(c->edit_remesh.end(), p, p + 3);
						c->edit_remesh_flags.push_back((uint8_t)(((f & 2) ? HCM_REMESH_BLOCK_EDIT : 0u) | ((f & 4) ? HCM_REMESH_SUNLIGHT : 0u)));
					} else
						out->n_pending++;
				}
			}
	if (!derive_pos.empty())
		if (int rc = hcm_rederive_chunks(c, (uint32_t)(derive_pos.size() / 3), derive_pos.data()))
			return rc;
	out->n_sunlight = n_changes, out->sunlight = c->edit_sunlight.data();
	out->n_remesh = (uint32_t)(c->edit_remesh.size() / 3), out->remesh_pos = c->edit_remesh.data(), out->remesh_flags = c->edit_remesh_flags.data();
	return HCM_OK;
}

int hcm_sync(hcm_ctx *c) {
	if (!c)
		return HCM_ERR_INVALID;
	CU(cudaSetDevice(c->device));
	CU(cudaStreamSynchronize(c->stream));
	return tune_harvest(c);
}

static int submit(hcm_ctx *c, uint32_t n, const int16_t *pos, int force_all) {
	if (n > c->max_tasks)
		return fail(c, HCM_ERR_CAPACITY, "batch of %u tasks exceeds max_tasks %u", n, c->max_tasks);
	if (n && !pos)
		return fail(c, HCM_ERR_INVALID, "null positions");
	CU(cudaSetDevice(c->device));
	c->tasks.resize(n);
	for (uint32_t i = 0; i < n; ++i) {
		int cx, cy, cz;
		if (!in_window(c, pos + 3 * i, &cx, &cy, &cz))
			return fail(c, HCM_ERR_OUT_OF_WINDOW, "task (%d,%d,%d) is outside the window", pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
		// the 27-array of ChunkTaskData<kMesh>::Pop (ChunkMeshTask.cpp:18-34): every neighbour must exist
		if (cx < 1 || cy < 1 || cz < 1 || cx > c->w.wx - 2 || cy > c->w.wy - 2 || cz > c->w.wz - 2)
			return fail(c, HCM_ERR_NOT_RESIDENT, "task (%d,%d,%d) touches the window border: a neighbour cannot be resident", pos[3 * i], pos[3 * i + 1],
			            pos[3 * i + 2]);
		if (c->nbr_count[cell(c, cx, cy, cz)] != 27) {
			for (int dy = -1; dy <= 1; ++dy)
				for (int dz = -1; dz <= 1; ++dz)
					for (int dx = -1; dx <= 1; ++dx)
						if (c->slot_of[cell(c, cx + dx, cy + dy, cz + dz)] < 0)
							return fail(c, HCM_ERR_NOT_RESIDENT, "task (%d,%d,%d): neighbour (%d,%d,%d) is not resident", pos[3 * i], pos[3 * i + 1],
							            pos[3 * i + 2], pos[3 * i] + dx, pos[3 * i + 1] + dy, pos[3 * i + 2] + dz);
			return fail(c, HCM_ERR_STATE, "internal: neighbour count out of sync");
		}
		c->tasks[i] = TaskIn{(int16_t)cx, (int16_t)cy, (int16_t)cz, 0};
	}
	c->batch_n = n;
	c->batch_pending = true, c->batch_done = false, c->counts_valid = false;
	c->force_all = force_all;
	if (n) {
		const int b = c->task_ring_next;
		c->task_ring_next ^= 1;
		if (c->task_ring_used[b])
			CU(cudaEventSynchronize(c->task_ring_ev[b])); // the copy of two batches ago has left this buffer
		memcpy(c->h_task_ring[b], c->tasks.data(), (size_t)n * sizeof(TaskIn));
		CU(cudaMemcpyAsync(c->d_tasks, c->h_task_ring[b], (size_t)n * sizeof(TaskIn), cudaMemcpyHostToDevice, c->stream));
		CU(cudaEventRecord(c->task_ring_ev[b], c->stream));
		c->task_ring_used[b] = true;
	}
	c->stats.h2d_bytes += (uint64_t)n * sizeof(TaskIn);
	return launch_batch(c);
}

int hcm_mesh_batch(hcm_ctx *c, uint32_t n, const int16_t *pos) {
	if (!c)
		return HCM_ERR_INVALID;
	return submit(c, n, pos, 0);
}

int hcm_mesh_wait(hcm_ctx *c, hcm_batch_counts *out) {
	if (!c)
		return HCM_ERR_INVALID;
	CU(cudaSetDevice(c->device));
	if (int rc = wait_batch(c))
		return rc;
	if (out) {
		out->n_tasks = c->batch_n, out->n_meshes = c->h_heads->meshes;
		out->n_vertices = c->h_heads->vtx, out->n_indices = c->h_heads->idx;
		out->n_worked = c->h_heads->n_work, out->_pad = 0;
	}
	return HCM_OK;
}

int hcm_mesh_fetch(hcm_ctx *c, hcm_batch_view *out) {
	if (!c || !out)
		return HCM_ERR_INVALID;
	CU(cudaSetDevice(c->device));
	if (int rc = wait_batch(c))
		return rc;
	const BatchHeads h = *c->h_heads;
	uint64_t tcap = c->h_task_cap, mcap = c->h_mesh_cap;
	if (int rc = host_reserve(c, &c->h_task_out, &tcap, c->batch_n))
		return rc;
	if (int rc = host_reserve(c, &c->h_mesh_out, &mcap, h.meshes))
		return rc;
	c->h_task_cap = (uint32_t)tcap, c->h_mesh_cap = (uint32_t)mcap;
	if (int rc = host_reserve(c, &c->h_vtx, &c->h_vtx_cap, h.vtx))
		return rc;
	if (int rc = host_reserve(c, &c->h_idx, &c->h_idx_cap, h.idx))
		return rc;
	cudaStream_t st = c->stream;
	CU(cudaEventRecord(c->ev[EV_F0], st));
	if (c->batch_n)
		CU(cudaMemcpyAsync(c->h_task_out, c->d_task_out, (size_t)c->batch_n * sizeof(hcm_task_result), cudaMemcpyDeviceToHost, st));
	if (h.meshes)
		CU(cudaMemcpyAsync(c->h_mesh_out, c->d_mesh_out, (size_t)h.meshes * sizeof(hcm_mesh_desc), cudaMemcpyDeviceToHost, st));
	if (h.vtx)
		CU(cudaMemcpyAsync(c->h_vtx, c->d_vtx, h.vtx * 8, cudaMemcpyDeviceToHost, st));
	if (h.idx)
		CU(cudaMemcpyAsync(c->h_idx, c->d_idx, h.idx * 2, cudaMemcpyDeviceToHost, st));
	CU(cudaEventRecord(c->ev[EV_F1], st));
	CU(cudaStreamSynchronize(st));
	c->ev_fetch = true;
	c->stats.d2h_bytes += (uint64_t)c->batch_n * sizeof(hcm_task_result) + (uint64_t)h.meshes * sizeof(hcm_mesh_desc) + h.vtx * 8 + h.idx * 2;
	out->n_tasks = c->batch_n, out->n_meshes = h.meshes, out->n_vertices = h.vtx, out->n_indices = h.idx;
	out->tasks = c->h_task_out, out->meshes = c->h_mesh_out, out->vertices = c->h_vtx, out->indices = c->h_idx;
	c->counts_valid = true, c->sink_meshes = h.meshes;
	return HCM_OK;
}

// ---- device-side sink ------------------------------------------------------------------------------------------------
int hcm_mesh_counts(hcm_ctx *c, hcm_batch_view *out) {
	if (!c || !out)
		return HCM_ERR_INVALID;
	CU(cudaSetDevice(c->device));
	if (int rc = wait_batch(c))
		return rc;
	const BatchHeads h = *c->h_heads;
	uint64_t tcap = c->h_task_cap, mcap = c->h_mesh_cap;
	if (int rc = host_reserve(c, &c->h_task_out, &tcap, c->batch_n))
		return rc;
	if (int rc = host_reserve(c, &c->h_mesh_out, &mcap, h.meshes))
		return rc;
	c->h_task_cap = (uint32_t)tcap, c->h_mesh_cap = (uint32_t)mcap;
	cudaStream_t st = c->stream;
	if (c->batch_n)
		CU(cudaMemcpyAsync(c->h_task_out, c->d_task_out, (size_t)c->batch_n * sizeof(hcm_task_result), cudaMemcpyDeviceToHost, st));
	if (h.meshes)
		CU(cudaMemcpyAsync(c->h_mesh_out, c->d_mesh_out, (size_t)h.meshes * sizeof(hcm_mesh_desc), cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	c->stats.d2h_bytes += (uint64_t)c->batch_n * sizeof(hcm_task_result) + (uint64_t)h.meshes * sizeof(hcm_mesh_desc);
	out->n_tasks = c->batch_n, out->n_meshes = h.meshes, out->n_vertices = h.vtx, out->n_indices = h.idx;
	out->tasks = c->h_task_out, out->meshes = c->h_mesh_out, out->vertices = nullptr, out->indices = nullptr;
	c->counts_valid = true, c->sink_meshes = h.meshes;
	return HCM_OK;
}

int hcm_mesh_emit_into(hcm_ctx *c, uint32_t n, const hcm_mesh_placement *placements, void *dev_vertex_base, uint64_t vertex_capacity, void *dev_index_base,
                       uint64_t index_capacity) {
	if (!c)
		return HCM_ERR_INVALID;
	if (!c->counts_valid)
		return fail(c, HCM_ERR_STATE, "hcm_mesh_emit_into needs the descriptors of the last batch: call hcm_mesh_counts (or hcm_mesh_fetch) first");
	if (n == 0)
		return HCM_OK;
	if (!placements || !dev_vertex_base || !dev_index_base)
		return fail(c, HCM_ERR_INVALID, "null placements / device buffers");
	if (((size_t)dev_vertex_base & 7) || ((size_t)dev_index_base & 1))
		return fail(c, HCM_ERR_INVALID, "device buffers must be aligned to sizeof(Vertex) = 8 and sizeof(Index) = 2");
	CU(cudaSetDevice(c->device));
	const uint32_t n_meshes = c->sink_meshes;
	// every range inside its buffer, no two ranges overlapping (a mesh placed twice would be legal but is a caller bug more often than not)
	std::vector<std::pair<uint64_t, uint64_t>> vr(n), ir(n);
	for (uint32_t i = 0; i < n; ++i) {
		const hcm_mesh_placement &pl = placements[i];
		if (pl.mesh >= n_meshes)
			return fail(c, HCM_ERR_INVALID, "placement %u: mesh %u of %u", i, pl.mesh, n_meshes);
		const hcm_mesh_desc &d = c->h_mesh_out[pl.mesh];
		if (pl.vertex_offset > vertex_capacity || d.vertex_count > vertex_capacity - pl.vertex_offset)
			return fail(c, HCM_ERR_INVALID, "placement %u: %u vertices at %llu do not fit %llu", i, d.vertex_count, (unsigned long long)pl.vertex_offset,
			            (unsigned long long)vertex_capacity);
		if (pl.index_offset > index_capacity || d.index_count > index_capacity - pl.index_offset)
			return fail(c, HCM_ERR_INVALID, "placement %u: %u indices at %llu do not fit %llu", i, d.index_count, (unsigned long long)pl.index_offset,
			            (unsigned long long)index_capacity);
		vr[i] = {pl.vertex_offset, pl.vertex_offset + d.vertex_count}, ir[i] = {pl.index_offset, pl.index_offset + d.index_count};
	}
	for (auto *r : {&vr, &ir}) {
		std::sort(r->begin(), r->end());
		for (uint32_t i = 1; i < n; ++i)
			if ((*r)[i].first < (*r)[i - 1].second)
				return fail(c, HCM_ERR_INVALID, "placements overlap in the %s buffer ([%llu,%llu) and [%llu,%llu))", r == &vr ? "vertex" : "index",
				            (unsigned long long)(*r)[i - 1].first, (unsigned long long)(*r)[i - 1].second, (unsigned long long)(*r)[i].first,
				            (unsigned long long)(*r)[i].second);
	}
	if (n > c->placement_cap) {
		if (c->d_placements) {
			c->stats.device_bytes -= (size_t)c->placement_cap * sizeof(hcm_mesh_placement);
			cudaFree(c->d_placements);
		}
		c->d_placements = nullptr, c->placement_cap = 0;
		const uint32_t cap = std::max(4096u, n + n / 4);
		CU(dev_alloc(c, &c->d_placements, cap));
		c->placement_cap = cap;
	}
	cudaStream_t st = c->stream;
	CU(cudaMemcpyAsync(c->d_placements, placements, (size_t)n * sizeof(hcm_mesh_placement), cudaMemcpyHostToDevice, st));
	CU(cudaEventRecord(c->ev[EV_S0], st));
	launch_scatter(c->d_placements, n, c->d_mesh_out, reinterpret_cast<const unsigned long long *>(c->d_vtx), c->d_idx, dev_vertex_base, dev_index_base,
	               c->grid * 8, st);
	CU(cudaGetLastError());
	CU(cudaEventRecord(c->ev[EV_S1], st));
	CU(cudaStreamSynchronize(st)); // `placements` may be pageable; the caller may hand the buffers to its renderer when this returns
	c->stats.kernel_launches += 1;
	c->stats.h2d_bytes += (uint64_t)n * sizeof(hcm_mesh_placement);
	c->ev_sink = true;
	return HCM_OK;
}

// ---- pipelined upload + mesh + fetch --------------------------------------------------------------------------------
// The three stages of a world (re)load overlap: while the copy engine moves upload block b+1 to the device, the mesh
// kernel works on the tasks whose 27 chunks arrived with blocks <= b, and the second copy engine returns the meshes of
// earlier sub-batches.  Same kernels, same arenas (the allocation heads simply keep counting across sub-batches).

int hcm_upload_and_mesh(hcm_ctx *c, uint32_t n_chunks, const int16_t *pos, const uint16_t *blocks, const int8_t *heights, uint32_t n_tasks,
                        const int16_t *task_pos, hcm_batch_view *out) {
	if (!c || !out)
		return HCM_ERR_INVALID;
	if (n_chunks == 0) { // nothing to overlap
		if (int rc = submit(c, n_tasks, task_pos, 0))
			return rc;
		return hcm_mesh_fetch(c, out);
	}
	if (n_tasks > c->max_tasks)
		return fail(c, HCM_ERR_CAPACITY, "batch of %u tasks exceeds max_tasks %u", n_tasks, c->max_tasks);
	if (n_tasks && !task_pos)
		return fail(c, HCM_ERR_INVALID, "null positions");
	CU(cudaSetDevice(c->device));
	const int wx = c->w.wx, wy = c->w.wy, wz = c->w.wz;
	std::vector<TaskIn> tin(n_tasks);
	for (uint32_t i = 0; i < n_tasks; ++i) {
		int cx, cy, cz;
		const int16_t *tp = task_pos + 3 * i;
		if (!in_window(c, tp, &cx, &cy, &cz))
			return fail(c, HCM_ERR_OUT_OF_WINDOW, "task (%d,%d,%d) is outside the window", tp[0], tp[1], tp[2]);
		if (cx < 1 || cy < 1 || cz < 1 || cx > wx - 2 || cy > wy - 2 || cz > wz - 2)
			return fail(c, HCM_ERR_NOT_RESIDENT, "task (%d,%d,%d) touches the window border: a neighbour cannot be resident", tp[0], tp[1], tp[2]);
		tin[i] = TaskIn{(int16_t)cx, (int16_t)cy, (int16_t)cz, 0};
	}
	c->batch_pending = false, c->batch_done = false, c->counts_valid = false; // an unfetched earlier batch is dropped, like a second hcm_mesh_batch
	std::vector<uint32_t> sub_mesh_end; // per fetched sub-batch: end of its descriptor range

	std::vector<uint32_t> perm(n_tasks), bucket_start; // perm[k] = caller's index of the k-th task in sub-batch order
	uint32_t submitted = 0, n_sub = 0, fetched = 0;
	std::vector<std::pair<uint32_t, uint32_t>> sub_range;
	BatchHeads prev{};
	bool overflow = false;
	uint64_t d2h = 0;
	cudaStream_t st = c->stream, fs = c->fetch_stream;

	auto fetch_completed = [&](bool block) -> int {
		while (fetched < n_sub) {
			if (block)
				CU(cudaEventSynchronize(c->sub_events[fetched]));
			else {
				cudaError_t q = cudaEventQuery(c->sub_events[fetched]);
				if (q == cudaErrorNotReady) {
					(void)cudaGetLastError(); // an answer, not an error
					return HCM_OK;
				}
				CU(q);
			}
			const BatchHeads cur = c->h_heads_ring[fetched];
			const uint32_t t0 = sub_range[fetched].first, t1 = sub_range[fetched].second;
			if (cur.overflow)
				overflow = true;
			if (!overflow) {
				uint64_t mcap = c->h_mesh_cap;
				if (int rc = host_reserve_keep(c, &c->h_mesh_out, &mcap, cur.meshes, prev.meshes, fs))
					return rc;
				c->h_mesh_cap = (uint32_t)mcap;
				if (int rc = host_reserve_keep(c, &c->h_vtx, &c->h_vtx_cap, cur.vtx, prev.vtx, fs))
					return rc;
				if (int rc = host_reserve_keep(c, &c->h_idx, &c->h_idx_cap, cur.idx, prev.idx, fs))
					return rc;
				if (t1 > t0)
					CU(cudaMemcpyAsync(c->h_task_stage + t0, c->d_task_out + t0, (size_t)(t1 - t0) * sizeof(hcm_task_result), cudaMemcpyDeviceToHost, fs));
				if (cur.meshes > prev.meshes)
					CU(cudaMemcpyAsync(c->h_mesh_out + prev.meshes, c->d_mesh_out + prev.meshes, (size_t)(cur.meshes - prev.meshes) * sizeof(hcm_mesh_desc),
					                   cudaMemcpyDeviceToHost, fs));
				if (cur.vtx > prev.vtx)
					CU(cudaMemcpyAsync(c->h_vtx + prev.vtx, c->d_vtx + prev.vtx, (cur.vtx - prev.vtx) * 8, cudaMemcpyDeviceToHost, fs));
				if (cur.idx > prev.idx)
					CU(cudaMemcpyAsync(c->h_idx + prev.idx, c->d_idx + prev.idx, (cur.idx - prev.idx) * 2, cudaMemcpyDeviceToHost, fs));
				d2h += (uint64_t)(t1 - t0) * sizeof(hcm_task_result) + (uint64_t)(cur.meshes - prev.meshes) * sizeof(hcm_mesh_desc) + (cur.vtx - prev.vtx) * 8 +
				       (cur.idx - prev.idx) * 2 + sizeof(BatchHeads);
			}
			prev = cur;
			sub_mesh_end.push_back(overflow ? 0u : cur.meshes);
			++fetched;
		}
		return HCM_OK;
	};

	const bool trace = getenv("HCM_TRACE") != nullptr; // host-side timeline of the pipeline on stderr
	const auto T0 = std::chrono::steady_clock::now();
	auto ms_now = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - T0).count(); };
	double t_assigned = 0, t_enqueued = 0, t_kernels = 0;

	UploadHooks hooks;
	hooks.assigned = [&](const std::vector<UploadEntry> &entries, uint32_t n_blocks) -> int {
		// every neighbour of every task must be resident once this upload is through (ChunkMeshTask.cpp:18-34)
		for (uint32_t i = 0; i < n_tasks; ++i)
			if (c->nbr_count[cell(c, tin[i].cx, tin[i].cy, tin[i].cz)] != 27)
				return fail(c, HCM_ERR_NOT_RESIDENT, "task (%d,%d,%d): a neighbour is neither resident nor part of this upload", task_pos[3 * i], task_pos[3 * i + 1],
				            task_pos[3 * i + 2]);
		// block a task becomes ready with = latest upload block among its 27 chunks: separable 3x3x3 maximum
		std::vector<int32_t> &cb = c->cell_block;
		cb.assign(c->n_cells, -1);
		for (size_t i = 0; i < entries.size(); ++i)
			cb[cell(c, entries[i].cx, entries[i].cy, entries[i].cz)] = (int32_t)c->layout.of((uint32_t)i);
		std::vector<int32_t> tmp(c->n_cells);
		const size_t sx = 1, sz = (size_t)wx, sy = (size_t)wx * wz;
		auto pass = [&](const std::vector<int32_t> &a, std::vector<int32_t> &b, size_t stride, int extent, int which) {
			for (int y = 0; y < wy; ++y)
				for (int z = 0; z < wz; ++z)
					for (int x = 0; x < wx; ++x) {
						const size_t i = cell(c, x, y, z);
						const int k = which == 0 ? x : which == 1 ? z : y;
						int32_t m = a[i];
						if (k > 0)
							m = std::max(m, a[i - stride]);
						if (k < extent - 1)
							m = std::max(m, a[i + stride]);
						b[i] = m;
					}
		};
		pass(cb, tmp, sx, wx, 0), pass(tmp, cb, sz, wz, 1), pass(cb, tmp, sy, wy, 2);
		bucket_start.assign(n_blocks + 1, 0);
		for (uint32_t i = 0; i < n_tasks; ++i)
			++bucket_start[std::max(tmp[cell(c, tin[i].cx, tin[i].cy, tin[i].cz)], 0) + 1];
		for (uint32_t b = 0; b < n_blocks; ++b)
			bucket_start[b + 1] += bucket_start[b];
		std::vector<uint32_t> fill(bucket_start.begin(), bucket_start.end() - 1);
		c->tasks.resize(n_tasks);
		for (uint32_t i = 0; i < n_tasks; ++i) {
			const uint32_t k = fill[std::max(tmp[cell(c, tin[i].cx, tin[i].cy, tin[i].cz)], 0)]++;
			perm[k] = i, c->tasks[k] = tin[i];
		}
		while (c->sub_events.size() < n_blocks) {
			cudaEvent_t e;
			CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
			c->sub_events.push_back(e);
		}
		if (c->heads_ring_cap < n_blocks) {
			if (c->h_heads_ring)
				cudaFreeHost(c->h_heads_ring);
			c->h_heads_ring = nullptr, c->heads_ring_cap = 0;
			CU(cudaHostAlloc((void **)&c->h_heads_ring, (size_t)n_blocks * sizeof(BatchHeads), cudaHostAllocDefault));
			c->heads_ring_cap = n_blocks;
		}
		if (int rc = host_reserve(c, &c->h_task_stage, &c->h_task_stage_cap, n_tasks))
			return rc;
		uint64_t tcap = c->h_task_cap;
		if (int rc = host_reserve(c, &c->h_task_out, &tcap, n_tasks))
			return rc;
		c->h_task_cap = (uint32_t)tcap;
		if (n_tasks)
			CU(cudaMemcpyAsync(c->d_tasks, c->tasks.data(), (size_t)n_tasks * sizeof(TaskIn), cudaMemcpyHostToDevice, st));
		CU(cudaMemsetAsync(c->d_heads, 0, sizeof(BatchHeads), st));
		c->stats.h2d_bytes += (uint64_t)n_tasks * sizeof(TaskIn);
		t_assigned = ms_now();
		return HCM_OK;
	};
	hooks.block_enqueued = [&](uint32_t blk, uint32_t n_blocks) -> int {
		const uint32_t ready = bucket_start[blk + 1];
		const bool last = blk + 1 == n_blocks;
		if (ready > submitted && (ready - submitted >= c->min_sub_batch || last)) {
			const uint32_t t0 = submitted, t1 = ready;
			CU(cudaMemsetAsync(&c->d_heads->n_work, 0, 2 * sizeof(unsigned int), st)); // n_work and cursor; the allocation heads keep counting
			launch_classify(c->w, c->d_tasks + t0, t1 - t0, c->d_worklist, c->d_heads, c->d_task_out + t0, 0, st);
			launch_mesh(c->maps, mesh_params(c, t0), c->grid, (int)c->tune_mesh.selected, st);
			CU(cudaGetLastError());
			CU(cudaMemcpyAsync(c->h_heads_ring + n_sub, c->d_heads, sizeof(BatchHeads), cudaMemcpyDeviceToHost, st));
			CU(cudaEventRecord(c->sub_events[n_sub], st));
			c->stats.kernel_launches += 2;
			sub_range.emplace_back(t0, t1);
			++n_sub;
			submitted = ready;
		}
		return fetch_completed(false);
	};
	if (int rc = upload_impl(c, n_chunks, pos, blocks, heights, nullptr, &hooks)) {
		cudaStreamSynchronize(st), cudaStreamSynchronize(fs);
		return rc;
	}
	t_enqueued = ms_now();
	if (int rc = fetch_completed(true))
		return rc;
	t_kernels = ms_now();
	CU(cudaStreamSynchronize(fs));
	if (trace)
		fprintf(stderr, "[hcm] upload_and_mesh: %u chunks, %u tasks, %u sub-batches | tasks sorted %.2f ms, uploads enqueued %.2f, last kernel done %.2f, last copy done %.2f\n",
		        n_chunks, n_tasks, n_sub, t_assigned, t_enqueued, t_kernels, ms_now());
	if (overflow) { // rare: grow the arenas like hcm_mesh_wait does and take the plain path; every chunk is resident by now
		CU(cudaStreamSynchronize(st));
		const uint64_t need_v = std::max<uint64_t>(prev.vtx, prev.idx / 6 * 4);
		if (int rc = alloc_arenas(c, std::max<uint64_t>(need_v + need_v / 8, c->vtx_cap), std::max<uint32_t>(prev.meshes + prev.meshes / 8 + 64, c->mesh_cap)))
			return rc;
		if (int rc = submit(c, n_tasks, task_pos, 0))
			return rc;
		return hcm_mesh_fetch(c, out);
	}
	for (uint32_t k = 0; k < n_tasks; ++k)
		c->h_task_out[perm[k]] = c->h_task_stage[k];
	// the kernel numbered each sub-batch's tasks from 0 in bucket order: hcm_mesh_desc::task is the caller's index
	for (uint32_t sb = 0, m0 = 0; sb < n_sub; m0 = sub_mesh_end[sb], ++sb)
		for (uint32_t m = m0; m < sub_mesh_end[sb]; ++m)
			c->h_mesh_out[m].task = perm[sub_range[sb].first + c->h_mesh_out[m].task];
	c->counts_valid = true, c->sink_meshes = prev.meshes; // hcm_mesh_emit_into may follow: the arenas hold the whole call's meshes
	c->stats.d2h_bytes += d2h;
	out->n_tasks = n_tasks, out->n_meshes = prev.meshes, out->n_vertices = prev.vtx, out->n_indices = prev.idx;
	out->tasks = c->h_task_out, out->meshes = c->h_mesh_out, out->vertices = c->h_vtx, out->indices = c->h_idx;
	return HCM_OK;
}

int hcm_get_stats(hcm_ctx *c, hcm_stats *out) {
	if (!c || !out)
		return HCM_ERR_INVALID;
	CU(cudaSetDevice(c->device));
	CU(cudaStreamSynchronize(c->stream));
	if (c->ev_upload) {
		CU(cudaEventElapsedTime(&c->stats.upload_copy_ms, c->ev[EV_UP0], c->ev[EV_UP1]));
		CU(cudaEventElapsedTime(&c->stats.derive_ms, c->ev[EV_UP1], c->ev[EV_UP2]));
	}
	if (c->ev_batch) {
		CU(cudaEventElapsedTime(&c->stats.classify_ms, c->ev[EV_B0], c->ev[c->batch_ab ? EV_BC : EV_B1]));
		CU(cudaEventElapsedTime(&c->stats.mesh_ms, c->ev[EV_B1], c->ev[EV_B2]));
	}
	if (c->ev_fetch)
		CU(cudaEventElapsedTime(&c->stats.fetch_copy_ms, c->ev[EV_F0], c->ev[EV_F1]));
	if (c->ev_sink)
		CU(cudaEventElapsedTime(&c->stats.sink_ms, c->ev[EV_S0], c->ev[EV_S1]));
	*out = c->stats;
	return HCM_OK;
}

void *hcm_stream(hcm_ctx *c) { return c ? (void *)c->stream : nullptr; }

int hcm_get_tune_info(hcm_ctx *c, hcm_tune_info *out) {
	if (!c || !out)
		return HCM_ERR_INVALID;
	CU(cudaSetDevice(c->device));
	CU(cudaStreamSynchronize(c->stream));
	if (int rc = tune_harvest(c))
		return rc;
	auto fill = [&](const KernelTune &t, hcm_kernel_tune &o) {
		o.selected = t.selected, o.rounds = t.rounds, o.n_variants = t.n_variants, o.baseline_ms = t.sum_base;
		for (int v = 0; v < HCM_MAX_VARIANTS; ++v)
			o.variant_ms[v] = t.sum_var[v], o.mismatches[v] = t.mismatches[v];
		o.state = t.forced >= 0 ? (uint32_t)HCM_TUNE_FORCED : ((!c->tune && t.state == HCM_TUNE_PENDING) ? (uint32_t)HCM_TUNE_OFF : t.state);
	};
	fill(c->tune_derive, out->derive), fill(c->tune_mesh, out->mesh);
	return HCM_OK;
}

int hcm_set_option(hcm_ctx *c, const char *name, int64_t value) {
	if (!c || !name)
		return HCM_ERR_INVALID;
	auto force = [&](KernelTune &t) -> int {
		if (value < -1 || value > (int64_t)t.n_variants)
			return fail(c, HCM_ERR_INVALID, "%s = %lld: 0 (baseline), 1 .. %u (variant) or -1 (qualification decides)", name, (long long)value, t.n_variants);
		t.forced = (int)value;
		if (value >= 0)
			t.selected = (uint32_t)value;
		else if (t.state != HCM_TUNE_VARIANT_ADOPTED)
			t.selected = 0;
		return HCM_OK;
	};
	const std::string n(name);
	if (n == "upload_adaptive")
		c->adaptive_upload = value != 0;
	else if (n == "tune")
		c->tune = value != 0;
	else if (n == "tune_fault")
		c->tune_fault = value != 0;
	else if (n == "derive_variant")
		return force(c->tune_derive);
	else if (n == "mesh_variant")
		return force(c->tune_mesh);
	else if (n == "tune_min_chunks" && value >= 1)
		c->tune_min_chunks = (uint32_t)std::min<int64_t>(value, 0x7fffffff);
	else if (n == "tune_min_tasks" && value >= 1)
		c->tune_min_tasks = (uint32_t)std::min<int64_t>(value, 0x7fffffff);
	else if (n == "tune_rounds" && value >= 1 && value <= 16)
		c->tune_rounds = (uint32_t)value;
	else
		return fail(c, HCM_ERR_INVALID, "hcm_set_option: unknown name or value out of range (%s = %lld)", name, (long long)value);
	return HCM_OK;
}

int hcm_debug_light(hcm_ctx *c, const int16_t pos[3], uint8_t *out) {
	if (!c || !pos || !out)
		return HCM_ERR_INVALID;
	CU(cudaSetDevice(c->device));
	if (!c->d_debug_light)
		CU(dev_alloc(c, &c->d_debug_light, (size_t)kLightBytes));
	uint8_t *dl = c->d_debug_light;
	int rc = submit(c, 1, pos, 1); // force the full path; task 0 -> debug slot 0
	if (!rc)
		rc = wait_batch(c);
	if (!rc) {
		cudaError_t e = cudaMemcpy(out, dl, kLightBytes, cudaMemcpyDeviceToHost);
		if (e != cudaSuccess)
			rc = fail(c, HCM_ERR_CUDA, "debug light copy failed: %s", cudaGetErrorString(e));
	}
	// leave debug mode
	c->d_debug_light = nullptr;
	cudaFree(dl);
	c->stats.device_bytes -= kLightBytes;
	c->batch_done = false, c->batch_pending = false;
	return rc;
}

#pragma GCC visibility pop
} // extern "C"
