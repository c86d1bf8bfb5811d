// Hand-written sm_100a kernels of the B200 chunk mesher.
//
//   derive_kernel   : per uploaded chunk -- raw Block data -> dense property-index grid, y-packed light planes,
//                     sunlight heights, chunk summary flags (streaming, HBM bound).
//   classify_kernel : per task -- chunks whose summary flags prove "no faces" are answered immediately; the rest
//                     go to a work list.
//   mesh_kernel     : persistent CTAs, one chunk at a time, everything on chip:
//                       TMA 3-D box load of the 36x34x36 halo'd voxel volume        (reference a2/a14: the gather)
//                       bit-plane face candidates -> exact BlockTexture::Show        (a8/a9)
//                       bit-parallel level-synchronous sunlight/torchlight flood     (a2-a4)
//                       per-face AO + smoothed light keys                            (a7)
//                       warp-parallel greedy merge in the reference's loop order     (a8)
//                       ordered compaction + vectorised vertex/index emission        (a12)
//                       custom and liquid-flow meshes                                (a6/a11/a13)
//
// (aN = row of SURVEY.md section 8a; reference file:line citations are at the individual functions.)
#include "hcm_device.cuh"

namespace hcm {

typedef unsigned long long u64;
typedef uint32_t u32;
constexpr u32 FULL = 0xffffffffu;

// ================================================================================================================
// small PTX helpers: mbarrier + TMA (cp.async.bulk.tensor)
// ================================================================================================================
#ifdef HCM_CUSIM // CPU test tier: the emulation's mbarrier / bulk-copy model (tests/cusim/cusim.cpp)
using cusim::mbar_expect_tx;
using cusim::mbar_init;
using cusim::mbar_wait;
__device__ __forceinline__ bool mbar_wait_bounded(u64 *bar, u32 phase) { return mbar_wait(bar, phase), true; }
using cusim::tma_load_1d;
using cusim::tma_load_3d;
__device__ __forceinline__ u32 lanemask_lt() { return (1u << cusim::self().lane) - 1u; }
__device__ __forceinline__ void fence_proxy_async() {}
#else
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 phase) {
	asm volatile("{\n"
	             ".reg .pred p;\n"
	             "WAIT_%=:\n"
	             "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
	             "@p bra DONE_%=;\n"
	             "bra WAIT_%=;\n"
	             "DONE_%=:\n"
	             "}" ::"r"(smem_u32(bar)),
	             "r"(phase)
	             : "memory");
}
// The same wait with a way out: the variant form of a kernel (hcm.h, "Kernel variants") must never be able to hang the context that
// is trying it out.  After ~0.25 s without the phase completing the wait gives up; the caller then works on whatever the buffer
// holds, the A/B round's hashes differ and the variant is rejected -- the baseline form re-writes the data in the same round.
__device__ __forceinline__ bool mbar_wait_bounded(u64 *bar, u32 phase) {
	u32 done = 0;
	unsigned long long t0 = 0;
	for (u32 polls = 0;; ++polls) {
		asm volatile("{\n"
		             ".reg .pred p;\n"
		             "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
		             "selp.u32 %0, 1, 0, p;\n"
		             "}"
		             : "=r"(done)
		             : "r"(smem_u32(bar)), "r"(phase)
		             : "memory");
		if (done)
			return true;
		if ((polls & 63u) == 63u) {
			unsigned long long now;
			asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
			if (t0 == 0)
				t0 = now;
			else if (now - t0 > 250000000ull)
				return false;
		}
	}
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, u64 *bar) {
	asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
	                 smem_u32(dst)),
	             "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
	             : "memory");
}

// 1-D TMA bulk copy global -> shared (size multiple of 16, both addresses 16-byte aligned)
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, u32 bytes, u64 *bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
	             "r"(smem_u32(bar))
	             : "memory");
}

__device__ __forceinline__ u32 lanemask_lt() {
	u32 m;
	asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
	return m;
}
// generic-proxy accesses to shared memory are ordered before the async-proxy (TMA) writes that follow
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif
__device__ __forceinline__ u32 warp_excl_scan(u32 v, int lane) {
	u32 x = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		u32 y = __shfl_up_sync(FULL, x, d);
		if (lane >= d)
			x += y;
	}
	return x - v;
}

#ifdef HCM_PHASE_TIMING // developer build: cycles per phase, summed over tasks (tools/phase_timing.py)
__device__ unsigned long long g_phase_cycles[64];
#define PHASE_MARK(i)                                                                                                  \
	do {                                                                                                               \
		if (tid == 0) {                                                                                                \
			const long long now_ = clock64();                                                                          \
			atomicAdd(&g_phase_cycles[i], (unsigned long long)(now_ - t_phase));                                       \
			t_phase = now_;                                                                                            \
		}                                                                                                              \
	} while (0)
#define PHASE_COUNT(i, n)                                                                                              \
	do {                                                                                                               \
		atomicAdd(&g_phase_cycles[i], (unsigned long long)(n));                                                        \
	} while (0)
#else
#define PHASE_COUNT(i, n)                                                                                              \
	do {                                                                                                               \
	} while (0)
#define PHASE_MARK(i)                                                                                                  \
	do {                                                                                                               \
	} while (0)
#endif

// ================================================================================================================
// derive kernel: one CTA per uploaded chunk, 256 threads; thread t owns columns x4..x4+3 of row z
// ================================================================================================================
// kForm 0 (baseline): the chunk's 64 KB arrive as ONE bulk copy; nothing is computed before the last byte has landed.
// The variant forms re-schedule that (same arithmetic, same bytes written; which one a context launches is decided on the device:
// hcm.h, "Kernel variants"):
//   quarters (forms 1, 3): FOUR 16 KB bulk copies (eight y-layers each) on four mbarriers.  The uniform test runs quarter by quarter
//     and a chunk that is not uniform leaves it at the first quarter that says so: its per-voxel lookups then run under the copies of
//     the later quarters, each waited for when the walk reaches its first layer.
//   hoist (forms 1, 2, 3): the sunlight-height load (1 KB of HBM, a full DRAM round trip that the baseline takes between the copy's
//     completion and the block-wide vote) is issued before the first wait.
//   L2 prefetch (forms 2, 3): the bytes in flight of the baseline are bounded by shared memory (3 CTAs x 64 KB per SM).  Thread 0
//     also reads the entry of the CTA one wave of resident CTAs further down the grid and, once its own copy has landed, asks the
//     L2 for that chunk's 64 KB + heights (cp.async.bulk.prefetch.L2): when that CTA starts, its copy is an L2 hit.
constexpr int kDeriveQuarterBytes = HCM_CHUNK_VOXELS * 2 / 4; // 16384: y-layers 8q .. 8q+7
constexpr int kDerivePrefetchDistance = 3 * 148;              // CTAs resident on a B200 at 3 per SM: one wave ahead
constexpr int kDeriveForms = 4;
#ifdef HCM_CUSIM
__device__ __forceinline__ void prefetch_l2_bulk(const void *, u32) {}
#else
__device__ __forceinline__ void prefetch_l2_bulk(const void *src, u32 bytes) {
	asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
#endif
// the entry one wave ahead, held by thread 0 (forms with the L2 prefetch only: the other forms carry nothing)
template <bool kOn> struct DeriveAhead {};
template <> struct DeriveAhead<true> {
	int32_t slot = -1;
	u32 flags = 0;
	__device__ __forceinline__ void load(const UploadEntry *entries, int tid) {
		if (tid == 0 && blockIdx.x + kDerivePrefetchDistance < gridDim.x) {
			const UploadEntry a = entries[blockIdx.x + kDerivePrefetchDistance];
			slot = a.slot, flags = a.flags;
		}
	}
	__device__ __forceinline__ void issue(const DevWorld &w, const int8_t *raw_heights, int tid) const { // after the CTA's own copy has landed
		if (tid == 0 && slot >= 0) {
			if (!(flags & UE_UNIFORM))
				prefetch_l2_bulk(w.raw + (size_t)slot * HCM_CHUNK_VOXELS, HCM_CHUNK_VOXELS * 2);
			prefetch_l2_bulk(raw_heights + (size_t)slot * HCM_CHUNK_COLUMNS, HCM_CHUNK_COLUMNS);
		}
	}
};
// kTrial: the launch is part of an A/B round.  Only there are the variant forms' waits bounded (a form being tried out must not be able
// to hang the context: after a time-out the CTA stops waiting, its output differs, the form is rejected); a form that was adopted or
// forced waits like the baseline does -- a bounded wait that gave up spuriously (a context switched out for a quarter of a second)
// would corrupt data with nobody comparing.
template <int kForm, bool kTrial = false>
__global__ void __launch_bounds__(256, 3) derive_kernel(DevWorld w, const UploadEntry *__restrict__ entries, const uint8_t *__restrict__ prop_index,
                                                     const DevTables *__restrict__ tables, const int8_t *__restrict__ raw_heights) {
	HCM_DYNAMIC_SMEM(derive_smem); // the chunk's 64 KB of Block data, one TMA bulk copy
	__shared__ uint16_t s_pinfo[256];                       // pflags | light level << 8
	__shared__ u32 s_idinfo[256];                           // by block id: row | pflags << 8 | level << 16, bit 31: the row depends on meta
	__shared__ u32 s_and, s_or;
	constexpr int kVariant = (kForm == 1 || kForm == 3) ? 1 : 0; // quarters
	constexpr bool kHoist = kForm != 0, kPrefetch = kForm >= 2;
	__shared__ __align__(8) u64 s_bars[kVariant == 0 ? 1 : 4];
	u64 &s_bar = s_bars[0];
	const int tid = threadIdx.x;
	const UploadEntry e = entries[blockIdx.x];
	bool gave_up = false; // kTrial only
	auto wait_for = [&](u64 *bar) {
		if constexpr (kTrial) {
			if (!gave_up)
				gave_up = !mbar_wait_bounded(bar, 0);
		} else
			mbar_wait(bar, 0);
	};
	DeriveAhead<kPrefetch> ahead;
	if constexpr (kPrefetch)
		ahead.load(entries, tid);
	const bool host_uniform = e.flags & UE_UNIFORM;
	if (tid == 0) {
		mbar_init(&s_bar, 1);
		if (kVariant == 1)
			for (int q = 1; q < 4; ++q)
				mbar_init(&s_bars[kVariant == 1 ? q : 0], 1);
		s_and = 0xffffffffu, s_or = 0;
	}
	if (tid == 0 && !host_uniform) { // the copy first, the tables while it is in flight
		if (kVariant == 0) {
			mbar_expect_tx(&s_bar, HCM_CHUNK_VOXELS * 2);
			tma_load_1d(derive_smem, w.raw + (size_t)e.slot * HCM_CHUNK_VOXELS, HCM_CHUNK_VOXELS * 2, &s_bar);
		} else
			for (int q = 0; q < 4; ++q) {
				mbar_expect_tx(&s_bars[kVariant == 1 ? q : 0], kDeriveQuarterBytes);
				tma_load_1d(derive_smem + q * kDeriveQuarterBytes, w.raw + (size_t)e.slot * HCM_CHUNK_VOXELS + q * (kDeriveQuarterBytes / 2), kDeriveQuarterBytes,
				            &s_bars[kVariant == 1 ? q : 0]);
			}
	}
	char4 hv = make_char4(0, 0, 0, 0); // hoist: the chunk's sunlight heights, requested before anything is waited for
	if (kHoist)
		hv = *reinterpret_cast<const char4 *>(raw_heights + (size_t)e.slot * HCM_CHUNK_COLUMNS + (tid >> 3) * 32 + ((tid & 7) << 2));
	s_pinfo[tid] = (uint16_t)(tables->pflags[tid] | ((tables->pmisc[tid] >> 4) << 8));
	s_idinfo[tid] = reinterpret_cast<const u32 *>(prop_index + 65536)[tid];
	__syncthreads();
	const uint16_t *sraw = reinterpret_cast<const uint16_t *>(derive_smem);

	const int z = tid >> 3, x4 = (tid & 7) << 2;
	const uint16_t *raw = w.raw + (size_t)e.slot * HCM_CHUNK_VOXELS;
	const size_t gx = (size_t)e.cx * 32 + x4, gz = (size_t)e.cz * 32 + z;
	uint8_t *prow = w.prop8 + (((size_t)e.cy * 32) * w.Z + gz) * w.XP + gx + kXMargin; // + y * Z * XP
	const size_t pstride = (size_t)w.Z * w.XP;
	const size_t col = ((size_t)e.cy * w.Z + gz) * w.X + gx;
	const size_t plane = (size_t)w.wy * w.Z * w.X;
	const size_t ci = ((size_t)e.cy * w.wz + e.cz) * w.wx + e.cx;

	// ---- uniform chunks (all air, all stone: most of a load sphere) need one table lookup, not 32768 ----------------
	u32 d0, pat;
	bool same;
	if (host_uniform) {
		// the host skipped the 64 KB copy: materialise the raw slot here (the liquid generator reads raw meta of neighbours)
		d0 = e.value, pat = d0 | (d0 << 16), same = true;
		uint16_t *rw = const_cast<uint16_t *>(raw);
#pragma unroll 8
		for (int y = 0; y < 32; ++y)
			*reinterpret_cast<uint2 *>(rw + x4 + (y * 32 + z) * 32) = make_uint2(pat, pat);
	} else if (kVariant == 0) {
		if constexpr (kForm == 0)
			mbar_wait(&s_bar, 0);
		else
			wait_for(&s_bar);
		if constexpr (kPrefetch)
			ahead.issue(w, raw_heights, tid);
		d0 = sraw[0], pat = d0 | (d0 << 16);
		u32 acc_and = 0xffffffffu, acc_or = 0; // uniform <=> AND == OR == pattern
#pragma unroll
		for (int y = 0; y < 32; ++y) {
			const uint2 r = *reinterpret_cast<const uint2 *>(sraw + x4 + (y * 32 + z) * 32);
			acc_and &= r.x & r.y, acc_or |= r.x | r.y;
		}
		same = acc_and == pat && acc_or == pat;
	} else {
		// quarter by quarter; the CTA leaves at the first quarter in which some thread saw a second value (the vote is block-wide,
		// so every thread takes the same way out) -- the later quarters are then waited for by the walk below
		wait_for(&s_bars[0]);
		d0 = sraw[0], pat = d0 | (d0 << 16);
		u32 acc_and = 0xffffffffu, acc_or = 0;
		same = true;
#pragma unroll 1
		for (int q = 0; q < 4; ++q) {
			if (q)
				wait_for(&s_bars[kVariant == 1 ? q : 0]);
			if constexpr (kPrefetch)
				if (q == 3)
					ahead.issue(w, raw_heights, tid); // the uniform test got to the last quarter: the whole copy has landed
#pragma unroll
			for (int y = 8 * q; y < 8 * q + 8; ++y) {
				const uint2 r = *reinterpret_cast<const uint2 *>(sraw + x4 + (y * 32 + z) * 32);
				acc_and &= r.x & r.y, acc_or |= r.x | r.y;
			}
			same = acc_and == pat && acc_or == pat;
			if (q < 3 && __syncthreads_or(!same)) {
				same = false;
				break;
			}
		}
	}
	if (!kHoist)
		*reinterpret_cast<char4 *>(w.height + col) = *reinterpret_cast<const char4 *>(raw_heights + (size_t)e.slot * HCM_CHUNK_COLUMNS + z * 32 + x4);
	else
		*reinterpret_cast<char4 *>(w.height + col) = hv;
	if constexpr (kPrefetch)
		if (host_uniform)
			ahead.issue(w, raw_heights, tid); // nothing of its own to wait for
	if (__syncthreads_and(same)) {
		const u32 pr = __ldg(prop_index + d0);
		const u32 info = s_pinfo[pr], fl = info & 0xffu, lvl = info >> 8;
		const u32 p4 = pr * 0x01010101u;
#pragma unroll 8
		for (int y = 0; y < 32; ++y)
			*reinterpret_cast<u32 *>(prow + y * pstride) = p4;
		const u32 pw = (fl & PF_INDIRECT) ? 0xffffffffu : 0u;
		*reinterpret_cast<uint4 *>(w.pass + col) = make_uint4(pw, pw, pw, pw);
#pragma unroll
		for (int l = 0; l < kMaxLightLevels; ++l)
			if (l < w.n_levels) {
				const u32 tw = (lvl && w.levels[l] == lvl) ? 0xffffffffu : 0u;
				*reinterpret_cast<uint4 *>(w.torch + l * plane + col) = make_uint4(tw, tw, tw, tw);
			}
		if (tid == 0) {
			u32 cf = CF_RESIDENT;
			if (fl & PF_EMPTY)
				cf |= CF_ALL_EMPTY | 0x3fu * CF_FACE_EMPTY0;
			if (fl & PF_SOLID)
				cf |= CF_ALL_SOLID | 0x3fu * CF_FACE_SOLID0;
			if (fl & PF_CUSTOM)
				cf |= CF_HAS_CUSTOM;
			if (fl & PF_DYNAMIC)
				cf |= CF_HAS_DYNAMIC;
			if (lvl)
				cf |= CF_HAS_TORCH;
			w.cflags[ci] = cf;
			w.slot[ci] = e.slot;
		}
		return;
	}

	// ---- general chunk -----------------------------------------------------------------------------------------------
	// Per voxel: one shared-memory lookup by block id (the 64 K table in global memory only for ids whose row depends on
	// meta), then bit inserts into the column words: pass (indirect light), sol / emp (class O / E).  The boundary-layer
	// summaries come out of those words after the loop.
	u32 pass[4] = {0, 0, 0, 0}, sol[4] = {0, 0, 0, 0}, emp[4] = {0, 0, 0, 0};
	u32 f_and = 0xffffffffu, f_or = 0; // info words accumulated over the thread's voxels: flags in bits 8..15, light level from bit 16
#pragma unroll 8
	for (int y = 0; y < 32; ++y) {
		if (kVariant == 1 && (y & 7) == 0 && y) { // the quarter this layer opens (already there if the uniform test got that far)
			wait_for(&s_bars[kVariant == 1 ? (y >> 3) : 0]);
			if constexpr (kPrefetch)
				if (y == 24)
					ahead.issue(w, raw_heights, tid); // (a second request for the same lines, if the uniform test already made it, is answered by the L2)
		}
		const uint2 r = *reinterpret_cast<const uint2 *>(sraw + x4 + (y * 32 + z) * 32);
		const u32 d[4] = {r.x & 0xffffu, r.x >> 16, r.y & 0xffffu, r.y >> 16};
		u32 info[4];
#pragma unroll
		for (int i = 0; i < 4; ++i)
			info[i] = s_idinfo[d[i] & 0xffu];
#pragma unroll
		for (int i = 0; i < 4; ++i) {
			if (info[i] >> 31) {
				const u32 pr = __ldg(prop_index + d[i]);
				info[i] = pr | ((u32)s_pinfo[pr] << 8);
			}
			pass[i] |= ((info[i] >> 8) & 1u) << y;  // PF_INDIRECT == 1
			sol[i] |= ((info[i] >> 10) & 1u) << y;  // PF_SOLID == 4
			emp[i] |= ((info[i] >> 11) & 1u) << y;  // PF_EMPTY == 8
			f_and &= info[i], f_or |= info[i];
		}
		*reinterpret_cast<u32 *>(prow + y * pstride) = (info[0] & 0xffu) | ((info[1] & 0xffu) << 8) | ((info[2] & 0xffu) << 16) | (info[3] << 24);
	}
	*reinterpret_cast<uint4 *>(w.pass + col) = make_uint4(pass[0], pass[1], pass[2], pass[3]);
	// light sources are rare: their column words are built in a second walk, only by the threads that saw one
#pragma unroll 1
	for (int l = 0; l < w.n_levels; ++l) {
		u32 t[4] = {0, 0, 0, 0};
		if (f_or >> 16) {
			const u32 want = l == 0 ? w.levels[0] : (l == 1 ? w.levels[1] : (l == 2 ? w.levels[2] : w.levels[3]));
#pragma unroll 1
			for (int y = 0; y < 32; ++y) {
				const uint2 r = *reinterpret_cast<const uint2 *>(sraw + x4 + (y * 32 + z) * 32);
				const u32 d[4] = {r.x & 0xffffu, r.x >> 16, r.y & 0xffffu, r.y >> 16};
#pragma unroll
				for (int i = 0; i < 4; ++i) {
					u32 info = s_idinfo[d[i] & 0xffu];
					if (info >> 31)
						info = (u32)s_pinfo[__ldg(prop_index + d[i])] << 8;
					t[i] |= ((info >> 16) == want ? 1u : 0u) << y;
				}
			}
		}
		*reinterpret_cast<uint4 *>(w.torch + l * plane + col) = make_uint4(t[0], t[1], t[2], t[3]);
	}
	// boundary layers: side 0 +x (x=31), 1 -x (x=0), 2 +y, 3 -y, 4 +z, 5 -z: is every voxel of the layer this thread sees class O / E?
	auto sides = [&](const u32 (&c)[4]) -> u32 {
		const u32 all = c[0] & c[1] & c[2] & c[3];
		u32 f = 0x3fu;
		if (x4 == 28 && c[3] != 0xffffffffu)
			f &= ~1u;
		if (x4 == 0 && c[0] != 0xffffffffu)
			f &= ~2u;
		if (!(all >> 31))
			f &= ~4u;
		if (!(all & 1u))
			f &= ~8u;
		if (z == 31 && all != 0xffffffffu)
			f &= ~16u;
		if (z == 0 && all != 0xffffffffu)
			f &= ~32u;
		return f;
	};
	const u32 face_solid = sides(sol), face_empty = sides(emp);

	// block-wide AND / OR of the per-thread summaries
	u32 a = ((f_and >> 8) & 0xffu) | (face_solid << 8) | (face_empty << 16);
	u32 o = ((f_or >> 8) & 0xffu) | ((f_or >> 16) ? 0x100u : 0u); // bit 8: some voxel emits light
	a = __reduce_and_sync(FULL, a);
	o = __reduce_or_sync(FULL, o);
	if ((tid & 31) == 0) {
		atomicAnd(&s_and, a);
		atomicOr(&s_or, o);
	}
	__syncthreads();
	if (tid == 0) {
		const u32 A = s_and, O = s_or;
		u32 cf = CF_RESIDENT;
		if (A & PF_EMPTY)
			cf |= CF_ALL_EMPTY;
		if (A & PF_SOLID)
			cf |= CF_ALL_SOLID;
		if (O & PF_CUSTOM)
			cf |= CF_HAS_CUSTOM;
		if (O & PF_DYNAMIC)
			cf |= CF_HAS_DYNAMIC;
		if (O & 0x100u)
			cf |= CF_HAS_TORCH;
		cf |= ((A >> 8) & 0x3f) * CF_FACE_SOLID0;
		cf |= ((A >> 16) & 0x3f) * CF_FACE_EMPTY0;
		w.cflags[ci] = cf;
		w.slot[ci] = e.slot;
	}
}

// Reset the dense grids of removed chunks to "air, nothing resident".
__global__ void __launch_bounds__(256) clear_kernel(DevWorld w, const UploadEntry *__restrict__ entries) {
	const UploadEntry e = entries[blockIdx.x];
	const int tid = threadIdx.x, z = tid >> 3, x4 = (tid & 7) << 2;
	const size_t gx = (size_t)e.cx * 32 + x4, gz = (size_t)e.cz * 32 + z;
	for (int y = 0; y < 32; ++y)
		*reinterpret_cast<u32 *>(w.prop8 + (((size_t)e.cy * 32 + y) * w.Z + gz) * w.XP + gx + kXMargin) = 0;
	const size_t col = ((size_t)e.cy * w.Z + gz) * w.X + gx;
	*reinterpret_cast<uint4 *>(w.pass + col) = make_uint4(0, 0, 0, 0);
	const size_t plane = (size_t)w.wy * w.Z * w.X;
	for (int l = 0; l < w.n_levels; ++l)
		*reinterpret_cast<uint4 *>(w.torch + l * plane + col) = make_uint4(0, 0, 0, 0);
	*reinterpret_cast<u32 *>(w.height + col) = 0;
	if (tid == 0) {
		const size_t ci = ((size_t)e.cy * w.wz + e.cz) * w.wx + e.cx;
		w.cflags[ci] = 0;
		w.slot[ci] = -1;
	}
}

// ================================================================================================================
// block edits: Chunk::SetBlock on the resident Block data, then the sunlight-height chain of the edited columns
// ================================================================================================================
// result word per edit: old block | changed << 16 | not resident << 17
__global__ void __launch_bounds__(128) edit_blocks_kernel(DevWorld w, const hcm_block_edit *__restrict__ edits, u32 n, u32 *__restrict__ result) {
	const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	const hcm_block_edit e = edits[i];
	const int cx = e.chunk[0] - w.ox, cy = e.chunk[1] - w.oy, cz = e.chunk[2] - w.oz;
	const int slot = w.slot[((size_t)cy * w.wz + cz) * w.wx + cx];
	if (slot < 0) {
		result[i] = 1u << 17;
		return;
	}
	uint16_t *raw = const_cast<uint16_t *>(w.raw) + (size_t)slot * HCM_CHUNK_VOXELS;
	const u32 old = raw[e.index & (HCM_CHUNK_VOXELS - 1)];
	const bool changed = old != e.block; // ChunkSetBlockTask.cpp:89-92
	if (changed)
		raw[e.index & (HCM_CHUNK_VOXELS - 1)] = e.block;
	result[i] = old | ((changed ? 1u : 0u) << 16);
}

// One thread per edited world column (cx, cz, x, z).  The host sorts the edits by column, highest chunk first, so the
// thread runs the chain of every changed edit of its column from top to bottom, one after the other: kFloodSunlight
// (ChunkFloodSunlightTask.cpp:41-57) gives the new height of a chunk's column, kSetSunlight (ChunkSetSunlightTask.cpp:
// 55-63) stores it and, if it changed, floods the chunk below.  Columns are independent of each other.
__global__ void __launch_bounds__(128) sunlight_chain_kernel(DevWorld w, const hcm_block_edit *__restrict__ edits, const uint2 *__restrict__ columns,
                                                             u32 n_columns, const u32 *__restrict__ result, const uint8_t *__restrict__ prop_index,
                                                             const DevTables *__restrict__ tables, int8_t *__restrict__ raw_heights,
                                                             hcm_sunlight_change *__restrict__ changes, u32 *__restrict__ n_changes, u32 cap) {
	const u32 c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= n_columns)
		return;
	const uint2 seg = columns[c]; // first edit, number of edits
	int visited_down_to = 0x7fffffff; // a chain that ran through a chunk already settled it
	for (u32 i = seg.x; i < seg.x + seg.y; ++i) {
		if (!((result[i] >> 16) & 1u))
			continue; // ChunkSetBlockTask.cpp:89-90: an unchanged block floods nothing
		const hcm_block_edit e = edits[i];
		const int cx = e.chunk[0] - w.ox, cz = e.chunk[2] - w.oz;
		const u32 x = e.index & 31u, z = (e.index >> 5) & 31u, xz = x + z * 32u;
		int cy = e.chunk[1] - w.oy;
		if (cy >= visited_down_to)
			continue;
		for (; cy >= 0 && cy + 1 < w.wy; --cy) {
			const int slot = w.slot[((size_t)cy * w.wz + cz) * w.wx + cx], up = w.slot[((size_t)(cy + 1) * w.wz + cz) * w.wx + cx];
			if (slot < 0 || up < 0)
				break; // Pop needs both chunks (ChunkFloodSunlightTask.cpp:20-23)
			int nh = HCM_CHUNK_SIZE;
			if (raw_heights[(size_t)up * HCM_CHUNK_COLUMNS + xz] <= 0) {
				const uint16_t *raw = w.raw + (size_t)slot * HCM_CHUNK_VOXELS;
				int y = HCM_CHUNK_SIZE - 1;
				while (y >= 0 && (tables->pmisc[prop_index[raw[x + (y * 32 + z) * 32]]] & 2u))
					--y;
				nh = y + 1;
			}
			int8_t *hp = raw_heights + (size_t)slot * HCM_CHUNK_COLUMNS + xz;
			const int old = *hp;
			visited_down_to = cy;
			if (nh == old)
				break;
			*hp = (int8_t)nh;
			const u32 k = atomicAdd(n_changes, 1u);
			if (k < cap) {
				hcm_sunlight_change ch{};
				ch.chunk[0] = e.chunk[0], ch.chunk[1] = (int16_t)(cy + w.oy), ch.chunk[2] = e.chunk[2];
				ch.index = (uint16_t)xz, ch.old_height = (int8_t)old, ch.new_height = (int8_t)nh;
				changes[k] = ch;
			}
		}
	}
}

// ================================================================================================================
// classify kernel: thread per task
// ================================================================================================================
__global__ void __launch_bounds__(256)
    classify_kernel(DevWorld w, const TaskIn *__restrict__ tasks, u32 n, u32 *__restrict__ worklist, BatchHeads *heads, hcm_task_result *task_out, int force_all) {
	const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n)
		return;
	const TaskIn ti = tasks[t];
	auto cf = [&](int dx, int dy, int dz) { return w.cflags[((size_t)(ti.cy + dy) * w.wz + (ti.cz + dz)) * w.wx + (ti.cx + dx)]; };
	const u32 c = cf(0, 0, 0);
	bool trivial = false;
	if ((c & CF_ALL_EMPTY) && !(c & CF_HAS_CUSTOM))
		trivial = true; // no textured voxel and no custom mesh in the chunk: nothing can be emitted
	else if (c & CF_ALL_SOLID) {
		// every face is between two class-O voxels as long as the six touching neighbour layers are class O too
		trivial = (cf(1, 0, 0) & (CF_FACE_SOLID0 << 1)) && (cf(-1, 0, 0) & (CF_FACE_SOLID0 << 0)) && (cf(0, 1, 0) & (CF_FACE_SOLID0 << 3)) &&
		          (cf(0, -1, 0) & (CF_FACE_SOLID0 << 2)) && (cf(0, 0, 1) & (CF_FACE_SOLID0 << 5)) && (cf(0, 0, -1) & (CF_FACE_SOLID0 << 4));
	}
	if (trivial && !force_all) {
		task_out[t] = hcm_task_result{0u, 0u};
	} else {
		// bit 31 of the work item: some chunk of the 27 emits light (the mesh kernel then runs the torch flood as well)
		u32 any = 0;
		for (int dy = -1; dy <= 1; ++dy)
			for (int dz = -1; dz <= 1; ++dz)
				for (int dx = -1; dx <= 1; ++dx)
					any |= cf(dx, dy, dz);
		const u32 i = atomicAdd(&heads->n_work, 1u);
		worklist[i] = t | ((any & CF_HAS_TORCH) ? 0x80000000u : 0u);
	}
}

// ================================================================================================================
// mesh kernel
// ================================================================================================================

constexpr int SM_VOX = 0;
constexpr int SM_LIGHT = SM_VOX + kVoxBytes;          // 58752
constexpr int SM_F = SM_LIGHT + kLvolBytes;           // 100368
constexpr int kFS = 33;                               // row stride of F (padded: lanes that walk different masks hit different banks)
constexpr int SM_CPLANE = SM_F + kMasks * kFS * 4;
constexpr int SM_REGION = SM_CPLANE + kInnerCols * 8; // 132656
constexpr int kRegionBytes = kGroupCap * 12;          // 73728
constexpr int SM_SMALL = SM_REGION + kRegionBytes;    // 206384
// The flood's inputs are staged by TMA in the phase region, which is idle from the end of a task's quad emission to the next
// task's plane phase: the boxes of task k+1 are fetched while task k emits its custom faces and descriptors.  The sunlight
// flood's own double buffer lies over the light volume, F and the custom-face plane (all written after the flood); the rarer
// torch flood, which must not touch the finished sunlight values, takes the phase region once the staged inputs are in registers.
constexpr int SM_FPASS = (SM_REGION + 127) / 128 * 128, SM_FHEIGHT = SM_FPASS + kFloodPassBytes;
static_assert(SM_FHEIGHT % 128 == 0, "TMA destinations");
static_assert(SM_FHEIGHT + kFloodHeightBytes <= SM_REGION + kRegionBytes, "flood input staging must fit the phase region");
constexpr int SM_SUNBUF = SM_LIGHT;

// region views
constexpr int RG_O = 0, RG_E = kInnerCols * 8, RG_D = 2 * kInnerCols * 8; // planes phase
constexpr int kSB = 64;                                                   // flood buffer stride (zero border)
constexpr int RG_SBUF = 0;                                                // 2 x 64 x 64 x 8 = 65536 (double buffered)
static_assert(2 * kSB * kSB * 8 <= kRegionBytes, "flood buffers must fit the phase region");
static_assert(3 * kInnerCols * 8 <= kRegionBytes, "planes must fit the phase region");
constexpr int RG_KEYS = 0, RG_GEO = kGroupCap * 8;
constexpr int RG_CLIST = 32768;                                           // exact-pair list behind the planes
constexpr int kCandCap = (kRegionBytes - RG_CLIST) / 4;
static_assert(3 * kInnerCols * 8 <= RG_CLIST, "planes overlap the candidate list");

struct Small {
	DevTables tab;
	u32 fc[kMasks + 2];    // visible faces per mask
	u32 fcpre[kMasks + 2]; // exclusive prefix
	u32 nq[kMasks + 2], nqo[kMasks + 2], nqt[kMasks + 2], baseo[kMasks + 2], baset[kMasks + 2];
	u32 qpre[kMasks + 2];                    // exclusive prefix of nq within the current group
	u32 n_quads, n_cand, cand_overflow;
	uint8_t order[3][kMasks + 2]; // non-empty masks of the current group: [0] heavy, [1] medium (dynamic warp scheduling), [2] tiny (one thread each)
	u32 n_order[3], next_mask;
	u32 rowmask[kMasks + 2]; // bit u: row u of the mask holds a face
	u32 crow_o[1024], crow_t[1024]; // custom faces per (y,z) row and stream -> exclusive prefix
	u32 aabb[2][kMaxSplits][6];
	u32 split_other[2][kMaxSplits];
	u32 wtmp[2][kWarps + 1];
	u64 bar, bar_flood;
	TaskIn next_task; // work item S.work_item: coordinates and worklist word, published by thread 0 a task ahead
	u32 next_item;
	u32 work_item, work_next, n_custom, any_dyn, Co, Ct, run_o, run_t, Qo, Qt, overflow, moff;
	u64 voff, ioff;
};
static_assert(SM_SUNBUF % 8 == 0 && SM_SUNBUF + 2 * kSB * kSB * 8 <= SM_REGION, "sunlight flood buffers end before the phase region");
constexpr int kSmemBytes = SM_SMALL + (int)sizeof(Small);
static_assert(kSmemBytes <= 232448, "shared memory budget (227 KB per CTA)");

__device__ __forceinline__ bool tex_show(const uint8_t *tf, u32 a, u32 b) { // BlockTexture::Show, BlockTexture.hpp:68-76
	const u32 fa = tf[a], fb = tf[b];
	if ((fa & TF_EMPTY) || a == b)
		return false;
	if (!(fa & HCM_TEX_TRANSPARENT) && !(fb & HCM_TEX_TRANSPARENT))
		return false;
	if ((fa & HCM_TEX_LIQUID) && !(fb & TF_EMPTY))
		return false;
	return !(fa & HCM_TEX_TRANSPARENT) || (fb & HCM_TEX_TRANSPARENT) || (fb & HCM_TEX_LIQUID);
}

__device__ __forceinline__ DynRec dyn_load(const DynRec *dyn, int x, int y, int z) {
	const u64 raw = __ldcg(reinterpret_cast<const u64 *>(dyn + cell_index(x, y, z)));
	DynRec r;
	*reinterpret_cast<u64 *>(&r) = raw;
	return r;
}

// get_tex of generate_regular_mesh_axis (BlockMeshAlgo.hpp:321-325): dynamic override first, else the property table
__device__ __forceinline__ u32 tex_code(const Small &S, const uint8_t *vox, const DynRec *dyn, int x, int y, int z, int face) {
	const u32 pr = vox[vox_index(x, y, z)];
	if (S.tab.pflags[pr] & PF_DYNAMIC) {
		const DynRec r = dyn_load(dyn, x, y, z);
		return ((r.texmask >> face) & 1) ? r.flow_code : 0u;
	}
	return S.tab.ptex[pr][face];
}

struct L4 { // Light4 of BlockMeshAlgo.hpp:18-46; byte v = corner v
	u32 sun, torch, ao;
};

// light4_init, BlockMeshAlgo.hpp:48-163.  (x,y,z) is the voxel owning the face; all 9 cells of the layer in front of
// the face are fetched once and shared by the four corners.
__device__ __forceinline__ void light4_corner(u32 fP, u32 fQ, u32 fD, bool dF, u32 lF, u32 lP, u32 lQ, u32 lD, u32 &sun, u32 &torch, u32 &ao) {
	const bool dP = fP & PF_DIRECT, dQ = fQ & PF_DIRECT, dD = fD & PF_DIRECT;
	ao = ((!dP && !dQ) || (!dD && !dF)) ? 0u : (u32)max(3 - (int)!dP - (int)!dD - (int)!dQ - (int)!dF, 0); // :127-131
	u32 cnt = 1, ss = lF & 15u, ts = lF >> 4; // :133-136
	if ((fP | fQ) & PF_INDIRECT) {           // :137-159
		if (fP & PF_INDIRECT)
			++cnt, ss += lP & 15u, ts += lP >> 4;
		if (fD & PF_INDIRECT)
			++cnt, ss += lD & 15u, ts += lD >> 4;
		if (fQ & PF_INDIRECT)
			++cnt, ss += lQ & 15u, ts += lQ >> 4;
	}
	ss <<= 2, ts <<= 2; // <= 240; exact floor division by 1..4 ((n*683)>>11 == n/3 for n < 1024, tests/test_host_logic.py)
	sun = cnt == 1 ? ss : (cnt == 2 ? ss >> 1 : (cnt == 3 ? (ss * 683u) >> 11 : ss >> 2)); // :160-161
	torch = cnt == 1 ? ts : (cnt == 2 ? ts >> 1 : (cnt == 3 ? (ts * 683u) >> 11 : ts >> 2));
}
__device__ __forceinline__ L4 light4_init(const Small &S, const uint8_t *vox, const uint8_t *light, int face, int x, int y, int z) {
	const int axis = face >> 1, dir = 1 - ((face & 1) << 1);
	constexpr int VY = kVZ * kVX, LX = kLY, LZ = kLY * kLS; // light volume strides: y 1, x 36, z 36*34
	// front cell (kLookup1v, :74) and the strides of the in-plane axes p < q
	const int vb = vox_index(x, y, z) + dir * (axis == 0 ? 1 : (axis == 1 ? VY : kVX));
	const int lb = lvol_index(x, y, z) + dir * (axis == 0 ? LX : (axis == 1 ? 1 : LZ));
	const int vp = axis == 0 ? VY : 1, vq = axis == 2 ? VY : kVX;
	const int lp = axis == 0 ? 1 : LX, lq = axis == 2 ? 1 : LZ;
	u32 fl[3][3], li[3][3];
#pragma unroll
	for (int dp = 0; dp < 3; ++dp)
#pragma unroll
		for (int dq = 0; dq < 3; ++dq) {
			fl[dp][dq] = S.tab.pflags[vox[vb + (dp - 1) * vp + (dq - 1) * vq]];
			li[dp][dq] = light[lb + (dp - 1) * lp + (dq - 1) * lq];
		}
	const bool dF = fl[1][1] & PF_DIRECT;
	// the four sign combinations along (p,q); kLookup3v (:75-108) orders them per face as
	//   faces +x,-y,+z : (-,-) (+,-) (+,+) (-,+)      faces -x,+y,-z : (-,-) (-,+) (+,+) (+,-)
	u32 s_mm, t_mm, a_mm, s_pm, t_pm, a_pm, s_pp, t_pp, a_pp, s_mp, t_mp, a_mp;
	light4_corner(fl[0][1], fl[1][0], fl[0][0], dF, li[1][1], li[0][1], li[1][0], li[0][0], s_mm, t_mm, a_mm);
	light4_corner(fl[2][1], fl[1][0], fl[2][0], dF, li[1][1], li[2][1], li[1][0], li[2][0], s_pm, t_pm, a_pm);
	light4_corner(fl[2][1], fl[1][2], fl[2][2], dF, li[1][1], li[2][1], li[1][2], li[2][2], s_pp, t_pp, a_pp);
	light4_corner(fl[0][1], fl[1][2], fl[0][2], dF, li[1][1], li[0][1], li[1][2], li[0][2], s_mp, t_mp, a_mp);
	const bool alt = (0x26u >> face) & 1; // faces 1, 2, 5
	const u32 s1 = alt ? s_mp : s_pm, t1 = alt ? t_mp : t_pm, a1 = alt ? a_mp : a_pm;
	const u32 s3 = alt ? s_pm : s_mp, t3 = alt ? t_pm : t_mp, a3 = alt ? a_pm : a_mp;
	L4 r;
	r.sun = s_mm | (s1 << 8) | (s_pp << 16) | (s3 << 24);
	r.torch = t_mm | (t1 << 8) | (t_pp << 16) | (t3 << 24);
	r.ao = a_mm | (a1 << 8) | (a_pp << 16) | (a3 << 24);
	return r;
}

// key of a regular face: 14 bits per corner (sun 6, torch 6, ao 2) + texture code in the top byte.
// Two faces merge iff texture and Light4 are equal (BlockMeshAlgo.hpp:364-378) iff their keys are equal.
__device__ __forceinline__ u64 pack_key(const L4 &l, u32 code) {
	u64 k = (u64)code << 56;
#pragma unroll
	for (int v = 0; v < 4; ++v) {
		const u64 c = ((l.sun >> (8 * v)) & 63u) | (((l.torch >> (8 * v)) & 63u) << 6) | (((l.ao >> (8 * v)) & 3u) << 12);
		k |= c << (14 * v);
	}
	return k;
}

__device__ __forceinline__ void aabb_atomic(Small &S, u32 s, u32 k, const u32 mn[3], const u32 mx[3]) {
#pragma unroll
	for (int c = 0; c < 3; ++c) {
		atomicMin(&S.aabb[s][k][c], mn[c]);
		atomicMax(&S.aabb[s][k][3 + c], mx[c]);
	}
}

// indices of one quad (BlockMeshAlgo.hpp:295-301 and :434-460) incl. the stale cur_vertex of a continuation mesh
__device__ __forceinline__ void store_indices(const MeshParams &p, u64 ioff, u32 qpos, u32 q, bool flip) {
	const u32 k = q / kQuadsPerMesh, j = q - k * kQuadsPerMesh;
	const u32 b = (j == 0 && k > 0) ? HCM_MESH_MAX_VERTICES : 4u * j;
	u32 i0, i1, i2, i3, i4, i5;
	if (flip)
		i0 = b, i1 = b + 1, i2 = b + 2, i3 = b, i4 = b + 2, i5 = b + 3;
	else
		i0 = b + 1, i1 = b + 2, i2 = b + 3, i3 = b, i4 = b + 1, i5 = b + 3;
	u32 *dst = reinterpret_cast<u32 *>(p.idx + ioff + 6ull * qpos);
	dst[0] = (i0 & 0xffffu) | (i1 << 16);
	dst[1] = (i2 & 0xffffu) | (i3 << 16);
	dst[2] = (i4 & 0xffffu) | (i5 << 16);
}

// ---------------------------------------------------------------------------------------------------------------
// liquid flow: restatement of flow_mesh_gen<Config,Variant>, block/block/public/Liquid.hpp:171-282
// ---------------------------------------------------------------------------------------------------------------
__device__ DynRec flow_record(const MeshParams &p, const Small &S, const uint8_t *vox, const TaskIn &ti, int x, int y, int z) {
	const u32 pr = vox[vox_index(x, y, z)];
	const DevMeshInfo mi = p.mesh_info[S.tab.pmesh[pr] - 1];
	const u32 unit = mi.height_unit;
	const int8_t nbh = (int8_t)(uint8_t)max(mi.variant * unit, 1u); // get_flow_height(Variant), Liquid.hpp:164-169,185
	int8_t nh[8];
	u32 liq = 0; // bit i: neighbour i is this liquid
	const int dxs[8] = {-1, 0, 1, -1, 1, -1, 0, 1}, dzs[8] = {-1, -1, -1, 0, 0, 1, 1, 1}; // Liquid.hpp:291-301
#pragma unroll
	for (int i = 0; i < 8; ++i) {
		const int nx = x + dxs[i], nz = z + dzs[i];
		const u32 q = vox[vox_index(nx, y, nz)];
		const u32 fl = S.tab.pflags[q];
		int8_t h;
		if (fl & PF_LIQ_SOURCE) {
			h = 16, liq |= 1u << i;
		} else if (fl & PF_LIQ_FLOW) {
			// raw meta, not the table-clamped variant (Liquid.hpp:181-182): fetch the uploaded Block data
			const int gx = ti.cx * 32 + nx, gy = ti.cy * 32 + y, gz = ti.cz * 32 + nz;
			const int32_t slot = p.w.slot[((size_t)(gy >> 5) * p.w.wz + (gz >> 5)) * p.w.wx + (gx >> 5)];
			const u32 data = p.w.raw[(size_t)slot * HCM_CHUNK_VOXELS + (gx & 31) + (((gy & 31) * 32 + (gz & 31)) * 32)];
			h = (int8_t)(uint8_t)max((data >> 8) * unit, 1u);
			liq |= 1u << i;
		} else
			h = (S.tab.pmisc[q] & 1) ? -1 : 0;
		nh[i] = h;
	}
	auto vh = [&](int a, int b, int c) -> uint8_t { // get_vh, Liquid.hpp:198-203
		const uint8_t mx = (uint8_t)max(max((int)nbh, (int)nh[a]), max((int)nh[b], (int)nh[c]));
		if (mx != 16 && ((nbh == 0) + (nh[a] == 0) + (nh[b] == 0) + (nh[c] == 0)) >= 3)
			return 0;
		return mx;
	};
	DynRec r;
	const uint8_t h0 = vh(0, 1, 3), h1 = vh(1, 2, 4), h2 = vh(3, 5, 6), h3 = vh(4, 6, 7);
	r.h[0] = h0, r.h[1] = h1, r.h[2] = h2, r.h[3] = h3;
	r.flow_code = mi.flow_code, r.flow_trans = mi.flow_trans;
	if (h0 == 16 && h1 == 16 && h2 == 16 && h3 == 16) { // Liquid.hpp:210-212
		r.texmask = 0x3f, r.faces = 0;
		return r;
	}
	const bool lf = h0 == 16 && h2 == 16, le = h0 == 0 && h2 == 0, rf = h1 == 16 && h3 == 16, re = h1 == 0 && h3 == 0;
	const bool ff = h2 == 16 && h3 == 16, fe = h2 == 0 && h3 == 0, bf = h0 == 16 && h1 == 16, be = h0 == 0 && h1 == 0;
	u32 faces = 1; // top
	if (!lf && !le && !((liq >> 3) & 1))
		faces |= 2;
	if (!rf && !re && !((liq >> 4) & 1))
		faces |= 4;
	if (!bf && !be && !((liq >> 1) & 1))
		faces |= 8;
	if (!ff && !fe && !((liq >> 6) & 1))
		faces |= 16;
	r.faces = (uint8_t)faces;
	// textures: right, left, top(none), bottom(flow), front, back -- Liquid.hpp:287-292
	r.texmask = (uint8_t)((rf ? 1u : 0u) | (lf ? 2u : 0u) | 8u | (ff ? 16u : 0u) | (bf ? 32u : 0u));
	return r;
}

// one face of a custom/dynamic mesh as the emitter consumes it
struct CFace {
	u32 axis, light_face, render_face, tex16;
	uint8_t v[4][4];
};
__device__ __forceinline__ void flow_face(const DynRec &r, u32 tex16, int which, CFace &f) { // Liquid.hpp:214-286
	const uint8_t h0 = r.h[0], h1 = r.h[1], h2 = r.h[2], h3 = r.h[3];
	f.tex16 = tex16;
	auto set = [&](int i, uint8_t a, uint8_t b, uint8_t c) { f.v[i][0] = a, f.v[i][1] = b, f.v[i][2] = c, f.v[i][3] = 4; };
	switch (which) {
	case 0: f.axis = 1, f.light_face = f.render_face = 2; set(0, 0, h0, 0), set(1, 0, h2, 16), set(2, 16, h3, 16), set(3, 16, h1, 0); break;
	case 1: f.axis = 0, f.light_face = f.render_face = 1; set(0, 0, h0, 0), set(1, 0, 0, 0), set(2, 0, 0, 16), set(3, 0, h2, 16); break;
	case 2: f.axis = 0, f.light_face = f.render_face = 0; set(0, 16, 0, 0), set(1, 16, h1, 0), set(2, 16, h3, 16), set(3, 16, 0, 16); break;
	case 3: f.axis = 2, f.light_face = f.render_face = 5; set(0, 0, 0, 0), set(1, 0, h0, 0), set(2, 16, h1, 0), set(3, 16, 0, 0); break;
	default: f.axis = 2, f.light_face = f.render_face = 4; set(0, 0, h2, 16), set(1, 0, 0, 16), set(2, 16, 0, 16), set(3, 16, h3, 16); break;
	}
}

// light4_interpolate, BlockMeshAlgo.hpp:165-196 (u32 arithmetic, ceil division by 16^3)
__device__ __forceinline__ u32 lerp16(u32 a, u32 b, u32 t) { return a * (u32)((int)16 - (int)t) + b * t; }
__device__ __forceinline__ u32 tri_interp(u32 lo, u32 hi, u32 du, u32 dv, u32 dw) {
	auto c = [](u32 w, int i) { return (w >> (8 * i)) & 0xffu; };
	const u32 s = lerp16(lerp16(lerp16(c(lo, 0), c(lo, 1), du), lerp16(c(lo, 3), c(lo, 2), du), dv),
	                     lerp16(lerp16(c(hi, 0), c(hi, 1), du), lerp16(c(hi, 3), c(hi, 2), du), dv), dw);
	return (s >> 12) + ((s & 4095u) ? 1u : 0u);
}

// ---------------------------------------------------------------------------------------------------------------
// bit-parallel flood of one light channel (reference a2-a4: ChunkMeshTask.cpp:48-111 + BlockLightAlgo.hpp:44-73).
// Level sets G_l = {cells with light >= l}: G_15 = sources(15), G_l = sources(l) | G_{l+1} | (N6(G_{l+1}) & passable).
// light(c) = 15 - (first step at which c enters G), step = 15 - l; the four bit planes of that value are kept in
// registers for the inner 34x34 columns and unpacked straight into the shared light volume.
// A column (x,z) of the 62^3 region is one 62-bit word (bit = y+15): y-neighbours are two shifts, x/z-neighbours four
// word loads from `sbuf`.  Each thread owns up to 3 inner + 6 outer columns.  Only columns within L1 distance
// 14-step of the inner square can still matter at `step` (the reference's IsBorderLightInterfere pruning), columns
// with nothing left to light are skipped, and the loop ends at the fixed point.
// ---------------------------------------------------------------------------------------------------------------
struct ColInfo {
	int c[kColsPerThread]; // sbuf index (with zero border) | last relevant step << 16, or -1 if unused
	u32 g0;                // element index in the dense planes of the lower chunk's word of domain column (0,0) (< 2^32: hcm_create)
	u32 X;                 // row pitch of the dense planes
	__device__ __forceinline__ int idx(int k) const { return c[k] & 0xffff; }
	__device__ __forceinline__ int lim(int k) const { return c[k] >> 16; }
	__device__ __forceinline__ u32 g(int k) const { // sbuf index = (z+1)*64 + (x+1)
		const int i = idx(k);
		return g0 + (u32)((i >> 6) - 1) * X + (u32)((i & 63) - 1);
	}
};

__device__ __forceinline__ u64 combine3(u32 lo, u32 mid, u32 hi) { // chunk words y-1, y, y+1 -> bits of y in [-15,46]
	return (u64)__funnelshift_r(lo, mid, 17) | ((u64)__funnelshift_r(mid, hi & 0x7fffu, 17) << 32); // two funnel shifts and a mask
}

// slot k of thread tid -> domain column (x,z) in [0,62)^2, or false
__device__ __forceinline__ bool slot_column(int tid, int k, int &x, int &z) {
	if (k < kInnerSlots) {
		const int ii = tid + k * kThreads;
		if (ii >= kInnerCols)
			return false;
		x = ii % kInner + 14, z = ii / kInner + 14;
		return true;
	}
	const int o = tid + (k - kInnerSlots) * kThreads;
	if (o < 14 * kFD) {
		z = o / kFD, x = o % kFD;
	} else if (o < 28 * kFD) {
		z = 48 + (o - 14 * kFD) / kFD, x = (o - 14 * kFD) % kFD;
	} else if (o < 28 * kFD + kInner * 28) {
		const int q = o - 28 * kFD;
		z = 14 + q / 28;
		const int xx = q % 28;
		x = xx < 14 ? xx : xx + kInner;
	} else
		return false;
	return true;
}

// spread the 4 low bits of n to bit 0 of 4 bytes
__device__ __forceinline__ u32 spread4(u32 n) { return (n * 0x00204081u) & 0x01010101u; }

template <bool kTorch>
__device__ __forceinline__ void flood_channel(const MeshParams &p, const ColInfo &ci, const u64 (&P)[kColsPerThread], u64 *sbuf, uint8_t *light, int tid) {
	const size_t plane_stride = (size_t)p.w.Z * p.w.X; // one chunk layer of columns
	u64 pl[kInnerSlots][4]; // bit planes of the light value of the inner columns
	// The level-set words live in `sbuf`, double buffered: step s reads buffer (s-1)&1 and publishes into buffer s&1, so
	// one barrier per step is enough.  A column that changed at step s is published at steps s and s+1 (once per buffer).
	u32 dirty = 0; // bit k: slot k changed in the previous step and still has to reach the other buffer
	u32 live = 0;  // bit k: slot k can still gain cells (valid, unlit passable cells left, still within reach of the inner square)
#pragma unroll
	for (int k = 0; k < kInnerSlots; ++k)
		pl[k][0] = pl[k][1] = pl[k][2] = pl[k][3] = 0;
	const int last_inject_step = kTorch ? 15 - (int)p.w.levels[p.w.n_levels - 1] : 0;
	bool any_set = false; // block-uniform: some G bit is set anywhere
#ifdef HCM_PHASE_TIMING
	long long tf0_ = clock64(), tfs_ = tf0_;
#endif
	for (int step = 0; step <= 14; ++step) {
		bool changed = false;
		const u32 val = 15 - step; // cells that enter G at this step get this light value
		const u64 *rd = sbuf + ((step - 1) & 1) * (kSB * kSB);
		u64 *wr = sbuf + (step & 1) * (kSB * kSB);
		u32 level_plane = ~0u; // torch: index of the source plane of this level, if any
		if (kTorch) {
			level_plane = ~0u;
			for (int l = 0; l < p.w.n_levels; ++l)
				if (p.w.levels[l] == val)
					level_plane = (u32)l;
		}
		const bool inject = kTorch ? level_plane != ~0u : step == 0;
		if (inject || (any_set && (live | dirty))) { // most threads have nothing left after a step or two: one test skips them
			u32 next_dirty = 0;
#pragma unroll
			for (int k = 0; k < kColsPerThread; ++k) {
				const bool lv = (live >> k) & 1, dt = (dirty >> k) & 1;
				if (ci.c[k] < 0 || !(inject || lv || dt))
					continue;
				const int c = ci.idx(k);
				const u64 g = step > 0 ? rd[c] : 0ull;
				u64 add = 0;
				PHASE_COUNT(27, 1); // column evaluations
				if (lv && step > 0) {
					if (step > ci.lim(k))
						live &= ~(1u << k);
					else {
						const u64 n = rd[c - 1] | rd[c + 1] | rd[c - kSB] | rd[c + kSB];
						add = (((g << 1) | (g >> 1) | n) & P[k]) & ~g;
					}
				}
				if (inject) { // sources of level 15 - step
					if (!kTorch) {
						add |= rd[c] & ~g; // sunlit cells, left there by the set-up pass (step 0: rd is the buffer nobody writes)
					} else {
						const u32 *src = p.w.torch + (size_t)level_plane * p.w.wy * plane_stride;
						const u32 gk = ci.g(k);
						add |= combine3(src[gk], src[gk + plane_stride], src[gk + 2 * plane_stride]) & ~g;
					}
				}
				const u64 ng = g | add;
				if (add) {
					PHASE_COUNT(25, 1); // ... that gained cells
					changed = true, next_dirty |= 1u << k;
					if (k < kInnerSlots) {
#pragma unroll
						for (int b = 0; b < 4; ++b)
							if ((val >> b) & 1)
								pl[k < kInnerSlots ? k : 0][b] |= add;
					}
				}
				if (add || dt)
					wr[c] = ng;
				if (inject || add) { // (re)evaluate whether the slot can still change
					if (P[k] & ~ng)
						live |= 1u << k;
					else
						live &= ~(1u << k);
				}
			}
			dirty = next_dirty;
		}
		const bool any_changed = __syncthreads_or(changed);
#ifdef HCM_PHASE_TIMING
		if (tid == 0) {
			PHASE_COUNT(28, 1); // flood steps executed
			const long long now_ = clock64();
			if (step == 0)
				PHASE_COUNT(31, now_ - tf0_); // step 0 (sources from global memory)
			else
				PHASE_COUNT(32, now_ - tfs_);
			tfs_ = now_;
		}
#endif
		any_set |= any_changed;
		if (!any_changed && step >= last_inject_step)
			break; // fixed point
	}
#ifdef HCM_PHASE_TIMING
	if (tid == 0) {
		PHASE_COUNT(29, clock64() - tf0_); // flood step loop
		tf0_ = clock64();
	}
#endif
	// unpack: 34 cells of a column = nine 32-bit words of the y-fastest light volume (cell y at byte y+1)
#pragma unroll
	for (int k = 0; k < kInnerSlots; ++k) {
		const int ii = tid + k * kThreads;
		if (ii < kInnerCols) {
			u32 *dst = reinterpret_cast<u32 *>(light + kLY * ii); // lvol_index(x,-1,z) = 36 * ((x+1) + 34*(z+1)) = 36 * ii
#pragma unroll
			for (int j = 0; j < 9; ++j) {
				// word j holds volume bytes 4j..4j+3 = cells y = 4j-1 .. 4j+2 = plane bits 4j+14 .. 4j+17
				const int sh = 4 * j + 14;
				u32 w = spread4((u32)(pl[k][0] >> sh) & 15u) | (spread4((u32)(pl[k][1] >> sh) & 15u) << 1) | (spread4((u32)(pl[k][2] >> sh) & 15u) << 2) |
				        (spread4((u32)(pl[k][3] >> sh) & 15u) << 3);
				if (kTorch)
					dst[j] |= w << 4;
				else
					dst[j] = w;
			}
		}
	}
	__syncthreads();
#ifdef HCM_PHASE_TIMING
	if (tid == 0)
		PHASE_COUNT(30, clock64() - tf0_); // unpack
#endif
}

// ---------------------------------------------------------------------------------------------------------------
// greedy merge of one mask by one warp, in the reference's scan order (BlockMeshAlgo.hpp:345-477).
// Lane = v (column).  Rows are visited in order; unconsumed equal-key neighbours form runs (width), a run grows
// downwards while every cell below matches (height), consumed cells are cleared from the row bit masks.
// Quad records replace the (already consumed) key slots of the mask, in emission order.
// ---------------------------------------------------------------------------------------------------------------
// geo word of a quad record: u | v<<5 | (width-1)<<10 | (height-1)<<15 | transparent<<20 | rank-in-stream-within-mask<<21
__device__ __forceinline__ u32 pack_geo(u32 u, u32 v, u32 w, u32 h, bool tr, u32 rank) {
	return u | (v << 5) | ((w - 1) << 10) | ((h - 1) << 15) | ((tr ? 1u : 0u) << 20) | (rank << 21);
}

// State lives in registers with lane = ROW: Frow (bit v: cell (u,v) is a face), EL (bit v: same key as the face to its
// left), ED (bit v: same key as the face below) and C (consumed cells).  Three passes:
//  1. set-up, lane = column: coalesced key loads, one ballot per EL/ED word, four rows in flight;
//  2. the sequential scan, one row per iteration, every lane doing the same word arithmetic on broadcast row words:
//       active cells   A = F[u] & ~C[u]
//       run starts     S = A & ~((A << 1) & EL[u])                                    (width,  BlockMeshAlgo.hpp:364-368)
//       run height     lane v owns the run that starts at column v; it stays alive at row r while ED[r-1] covers its
//                      cells; the union of the live runs (redux.or) is what row r loses        (:371-379, :462-466)
//     A cell below an unconsumed cell can only be consumed by the very quad that starts at that cell, so the height
//     test needs no activity check.  A warp issues one instruction every ~7 cycles here, so this pass is kept to the
//     bare recurrence: it only leaves (u, v, width, height) per quad, in emission order;
//  3. lane = quad: fetch the key of the quad's first cell, split into the opaque / transparent streams (ranks by
//     ballot) and finish the records.  Records overwrite the mask's (already read) key slots in emission order.
// kVariant 1 re-schedules the height test of pass 2 (hcm.h, "Kernel variants"): the row directly below a run is tested on its own
// (most runs end there), after that FOUR rows per round -- their ED words come from four independent shuffles and the four unions
// from four independent reductions, so a round of four rows waits for one shuffle and one reduction instead of four of each.  A run
// that has died stays dead (alive == 0 passes every later test as 0), and rows past 31 never match a lane, so the extra rows of
// the last round change nothing: the quads and the consumed sets are the baseline's.
template <int kVariant>
__device__ void merge_mask(Small &S, const u32 *F, u64 *keys, u32 *geo, int m, int g0, int lane) {
	const u32 lt = (1u << lane) - 1u;
	const u32 Frow = F[m * kFS + lane];
	const u32 pre = warp_excl_scan(__popc(Frow), lane);
	const u32 base = S.fcpre[m] - S.fcpre[g0];
	const u32 rows = __ballot_sync(FULL, Frow != 0);
	u32 EL = 0, ED = 0;
#ifdef HCM_PHASE_TIMING
	const long long tms0_ = clock64();
#endif
	{
		// a face key is never 0 (its texture code is not 0), so k == 0 stands for "no face here"
		u64 kprev = 0;
		int uprev = -2;
		u32 rr = rows;
		while (rr) {
			int u[4];
			u32 Fu[4], pu[4];
			u64 k[4];
#pragma unroll
			for (int j = 0; j < 4; ++j) {
				u[j] = rr ? __ffs(rr) - 1 : 32;
				rr &= rr - 1;
			}
#pragma unroll
			for (int j = 0; j < 4; ++j) {
				Fu[j] = __shfl_sync(FULL, Frow, u[j] & 31) & (u[j] < 32 ? 0xffffffffu : 0u);
				pu[j] = __shfl_sync(FULL, pre, u[j] & 31);
			}
#pragma unroll
			for (int j = 0; j < 4; ++j)
				k[j] = ((Fu[j] >> lane) & 1) ? keys[base + pu[j] + __popc(Fu[j] & lt)] : 0ull;
#pragma unroll
			for (int j = 0; j < 4; ++j) {
				u64 kl = __shfl_up_sync(FULL, k[j], 1);
				if (lane == 0)
					kl = 0;
				const u64 kup = j == 0 ? kprev : k[j ? j - 1 : 0];
				const int uup = j == 0 ? uprev : u[j ? j - 1 : 0];
				const u32 el = __ballot_sync(FULL, k[j] != 0 && k[j] == kl);
				const u32 ed = __ballot_sync(FULL, k[j] != 0 && k[j] == kup);
				if (lane == u[j])
					EL = el;
				if (uup == u[j] - 1 && lane == uup)
					ED = ed;
			}
			kprev = k[3], uprev = u[3];
		}
	}
#ifdef HCM_PHASE_TIMING
	const long long tms1_ = clock64();
	u32 iters_ = 0, hsteps_ = 0, hrounds_ = 0, hdeep_ = 0;
#endif
	u32 C = 0, nq = 0;
	const u32 right = ~((2u << lane) - 1u); // columns right of this lane
	u32 rr = rows;
	if constexpr (kVariant == 2) {
		// Two rows per iteration.  A row changes what a later row sees only through the cells its runs consume below it, and four rows
		// out of five consume nothing (every run one row high: profiles/r02_emulation_event_counts.txt).  So row u and the next non-empty
		// row u2 are scanned TOGETHER -- row words, run starts, run masks and the first height test of both, two independent chains in
		// one instruction stream -- with u2 reading the consumed set as it stood before u.  That reading is right unless u2 = u + 1 and
		// some run of u reaches into it (un1 != 0): then u2 is left for the next iteration, which sees the updated set.  (u2 > u + 1
		// means row u + 1 holds no face, so no run of u can extend at all.)  Quads come out in the same order: u's, then u2's.
		auto finish = [&](int u, u32 startmask, u32 rm, u32 alive, u32 un) { // first height test done: alive / un are its result
			u32 height = 1 + (alive != 0);
			int r = u + 1;
			if (lane == r)
				C |= un;
			++r;
			while (un) { // rare: some run is more than two rows high -- four rows per round, as in form 1
				const u32 e0 = __shfl_sync(FULL, ED, (r - 1) & 31), e1 = __shfl_sync(FULL, ED, r & 31);
				const u32 e2 = __shfl_sync(FULL, ED, (r + 1) & 31), e3 = __shfl_sync(FULL, ED, (r + 2) & 31);
				const u32 a0 = (e0 & alive) == alive ? alive : 0u;
				const u32 a1 = (e1 & a0) == a0 ? a0 : 0u;
				const u32 a2 = (e2 & a1) == a1 ? a1 : 0u;
				const u32 a3 = (e3 & a2) == a2 ? a2 : 0u;
				const u32 un0 = __reduce_or_sync(FULL, a0), un1 = __reduce_or_sync(FULL, a1);
				const u32 un2 = __reduce_or_sync(FULL, a2), un3 = __reduce_or_sync(FULL, a3);
				height += (a0 != 0) + (a1 != 0) + (a2 != 0) + (a3 != 0);
				const int d = lane - r;
				C |= d == 0 ? un0 : (d == 1 ? un1 : (d == 2 ? un2 : (d == 3 ? un3 : 0u)));
				alive = a3, un = un3, r += 4;
			}
			if (rm)
				geo[base + nq + __popc(startmask & lt)] = (u32)u | ((u32)lane << 5) | ((u32)(__popc(rm) - 1) << 10) | ((height - 1) << 15);
			nq += __popc(startmask);
		};
		while (rr) {
			const int u = __ffs(rr) - 1;
			rr &= rr - 1;
			const int u2 = rr ? __ffs(rr) - 1 : 32; // the partner row (still in rr)
			const u32 Fu = __shfl_sync(FULL, Frow, u), ELu = __shfl_sync(FULL, EL, u), Cu = __shfl_sync(FULL, C, u), e1 = __shfl_sync(FULL, ED, u);
			const u32 Fv = __shfl_sync(FULL, Frow, u2 & 31), ELv = __shfl_sync(FULL, EL, u2 & 31), Cv = __shfl_sync(FULL, C, u2 & 31),
			          e2 = __shfl_sync(FULL, ED, u2 & 31);
			const u32 A = Fu & ~Cu, B = u2 < 32 ? (Fv & ~Cv) : 0u;
			const u32 sm1 = A & ~((A << 1) & ELu), sm2 = B & ~((B << 1) & ELv);
			const u32 nb1 = (sm1 | ~A) & right, nb2 = (sm2 | ~B) & right;
			const u32 low1 = nb1 & (0u - nb1), low2 = nb2 & (0u - nb2);
			const u32 rm1 = ((sm1 >> lane) & 1) ? ((low1 - 1u) & ~lt) : 0u, rm2 = ((sm2 >> lane) & 1) ? ((low2 - 1u) & ~lt) : 0u;
			const u32 al1 = (e1 & rm1) == rm1 ? rm1 : 0u, al2 = (e2 & rm2) == rm2 ? rm2 : 0u;
			const u32 un1 = __reduce_or_sync(FULL, al1), un2 = __reduce_or_sync(FULL, al2);
#ifdef HCM_PHASE_TIMING
			++iters_;
#endif
			if (A)
				finish(u, sm1, rm1, al1, un1);
			if (u2 < 32 && !(u2 == u + 1 && un1)) { // the partner's reading of the consumed set was right
				rr &= rr - 1;
				if (B)
					finish(u2, sm2, rm2, al2, un2);
#ifdef HCM_PHASE_TIMING
				++hdeep_; // (form 2: rows that went through as partners)
#endif
			}
		}
	} else
	while (rr) {
		const int u = __ffs(rr) - 1;
		rr &= rr - 1;
		const u32 Fu = __shfl_sync(FULL, Frow, u), ELu = __shfl_sync(FULL, EL, u);
		const u32 A = Fu & ~__shfl_sync(FULL, C, u);
		if (!A)
			continue;
#ifdef HCM_PHASE_TIMING
		++iters_;
#endif
		const u32 startmask = A & ~((A << 1) & ELu); // unconsumed, and not continuing an equal unconsumed left neighbour
		const u32 nb = (startmask | ~A) & right;     // run boundaries right of this lane
		const u32 low = nb & (0u - nb);
		const bool start = (startmask >> lane) & 1;
		const u32 rm = start ? ((low - 1u) & ~lt) : 0u; // cells of the run that starts here
		u32 height = 1, alive = rm;
		if (kVariant == 0) {
			for (int r = u + 1;; ++r) { // ED[31] == 0 ends every run
#ifdef HCM_PHASE_TIMING
				++hsteps_, ++hrounds_;
#endif
				const u32 e = __shfl_sync(FULL, ED, (r - 1) & 31);
				alive = (e & alive) == alive ? alive : 0u;
				const u32 un = __reduce_or_sync(FULL, alive);
				if (!un)
					break;
				height += alive != 0;
				if (lane == r)
					C |= un;
			}
		} else {
			int r = u + 1; // the row whose cells the next test may consume; ED[31] == 0 ends every run at r = 32
			u32 un;
			{
#ifdef HCM_PHASE_TIMING
				++hsteps_, ++hrounds_;
#endif
				const u32 e = __shfl_sync(FULL, ED, (r - 1) & 31);
				alive = (e & alive) == alive ? alive : 0u;
				un = __reduce_or_sync(FULL, alive);
				height += alive != 0;
				if (lane == r)
					C |= un;
				++r;
			}
#ifdef HCM_PHASE_TIMING
			if (un)
				++hdeep_; // some run of this row reaches into the next one
#endif
			while (un) {
#ifdef HCM_PHASE_TIMING
				hsteps_ += 4, ++hrounds_;
#endif
				const u32 e0 = __shfl_sync(FULL, ED, (r - 1) & 31), e1 = __shfl_sync(FULL, ED, r & 31);
				const u32 e2 = __shfl_sync(FULL, ED, (r + 1) & 31), e3 = __shfl_sync(FULL, ED, (r + 2) & 31);
				const u32 a0 = (e0 & alive) == alive ? alive : 0u;
				const u32 a1 = (e1 & a0) == a0 ? a0 : 0u;
				const u32 a2 = (e2 & a1) == a1 ? a1 : 0u;
				const u32 a3 = (e3 & a2) == a2 ? a2 : 0u;
				const u32 un0 = __reduce_or_sync(FULL, a0), un1 = __reduce_or_sync(FULL, a1);
				const u32 un2 = __reduce_or_sync(FULL, a2), un3 = __reduce_or_sync(FULL, a3);
				height += (a0 != 0) + (a1 != 0) + (a2 != 0) + (a3 != 0);
				const int d = lane - r; // rows r .. r+3 lose what the runs still alive there cover
				C |= d == 0 ? un0 : (d == 1 ? un1 : (d == 2 ? un2 : (d == 3 ? un3 : 0u)));
				alive = a3, un = un3, r += 4;
			}
		}
		if (start)
			geo[base + nq + __popc(startmask & lt)] = (u32)u | ((u32)lane << 5) | ((u32)(__popc(rm) - 1) << 10) | ((height - 1) << 15);
		nq += __popc(startmask);
	}
#ifdef HCM_PHASE_TIMING
	const long long tms2_ = clock64();
#endif
	u32 nqt = 0;
	__syncwarp(); // pass 2 wrote the geometry words lane by lane; pass 3 reads them quad by quad
	for (u32 i0 = 0; i0 < nq; i0 += 32) {
		const u32 i = i0 + lane;
		const bool valid = i < nq;
		const u32 g = valid ? geo[base + i] : 0u;
		const u32 Fr = __shfl_sync(FULL, Frow, g & 31), pr = __shfl_sync(FULL, pre, g & 31);
		u64 key = 0;
		bool tr = false;
		if (valid) {
			key = keys[base + pr + __popc(Fr & ((1u << ((g >> 5) & 31)) - 1u))];
			tr = S.tab.texflags[key >> 56] & HCM_TEX_TRANS_PASS; // UseTransparentPass, BlockMeshAlgo.hpp:394
		}
		const u32 trm = __ballot_sync(FULL, tr);
		__syncwarp();
		if (valid) { // slot i <= position of the quad's first cell < position of any later quad's first cell
			const u32 tb = nqt + __popc(trm & lt);
			keys[base + i] = key;
			geo[base + i] = g | ((tr ? 1u : 0u) << 20) | ((tr ? tb : i - tb) << 21);
		}
		nqt += __popc(trm);
	}
#ifdef HCM_PHASE_TIMING
	if (lane == 0) {
		PHASE_COUNT(37, tms1_ - tms0_);       // set-up loop cycles
		PHASE_COUNT(38, tms2_ - tms1_);       // main loop cycles
		PHASE_COUNT(39, iters_);              // main loop iterations
		PHASE_COUNT(55, hsteps_);             // height steps (rows tested)
		PHASE_COUNT(57, hrounds_);            // ... in this many dependent shuffle -> reduction round trips
		PHASE_COUNT(58, hdeep_);              // (variant form) row iterations with a run taller than one row
		PHASE_COUNT(56, clock64() - tms2_);   // record pass cycles
	}
#endif
	if (lane == 0)
		S.nq[m] = nq, S.nqt[m] = nqt, S.nqo[m] = nq - nqt;
}

// The same scan done by ONE thread, for masks with a handful of faces (4 out of 5 masks of a terrain chunk have at most
// 16): 32 such masks run side by side in one warp.  Consumed cells are marked in place with a key no face can have.
constexpr u64 kConsumedKey = ~0ull; // sun field 63 > 60
#ifndef HCM_THREAD_MERGE_MAX
#define HCM_THREAD_MERGE_MAX 12
#endif
#ifndef HCM_TINY_PER_WARP
#define HCM_TINY_PER_WARP 8
#endif
constexpr u32 kThreadMergeMax = HCM_THREAD_MERGE_MAX; // faces
constexpr int kTinyPerWarp = HCM_TINY_PER_WARP;     // tiny masks per warp: fewer lanes per warp = less divergence, and spare warps are plentiful here
__device__ void merge_mask_thread(Small &S, const u32 *F, u64 *keys, u32 *geo, int m, int g0) {
	const u32 *Fm = F + m * kFS;
	const u32 base = S.fcpre[m] - S.fcpre[g0];
	u32 nq = 0, nqt = 0, rowslot = base;
	for (u32 rr = S.rowmask[m]; rr; rr &= rr - 1) { // non-empty rows only: a tiny mask has three of them on average
		const int u = __ffs(rr) - 1;
		const u32 Fu = Fm[u];
		u32 bits = Fu;
		while (bits) {
			const int v = __ffs(bits) - 1;
			const u32 idx = rowslot + __popc(Fu & ((1u << v) - 1u));
			const u64 key = keys[idx];
			if (key == kConsumedKey) {
				bits &= bits - 1;
				continue;
			}
			u32 w = 1; // BlockMeshAlgo.hpp:364-368 (consumed or missing cells never compare equal)
			while (v + w < 32 && ((Fu >> (v + w)) & 1u) && keys[idx + w] == key)
				++w;
			const u32 runmask = (w == 32 ? 0xffffffffu : ((1u << w) - 1u)) << v;
			u32 h = 1, ps = rowslot + __popc(Fu); // :371-379
			for (int uu = u + 1; uu < 32; ++uu) {
				const u32 Fr = Fm[uu];
				if ((Fr & runmask) != runmask)
					break;
				const u32 i2 = ps + __popc(Fr & ((1u << v) - 1u));
				bool ok = true;
				for (u32 k = 0; k < w && ok; ++k)
					ok = keys[i2 + k] == key;
				if (!ok)
					break;
				for (u32 k = 0; k < w; ++k)
					keys[i2 + k] = kConsumedKey; // :462-466
				++h, ps += __popc(Fr);
			}
			const bool tr = S.tab.texflags[key >> 56] & HCM_TEX_TRANS_PASS;
			keys[base + nq] = key; // record slot <= idx: never an unread cell
			geo[base + nq] = pack_geo((u32)u, (u32)v, w, h, tr, tr ? nqt : nq - nqt);
			++nq, nqt += tr ? 1u : 0u;
			bits &= ~runmask;
		}
		rowslot += __popc(Fu);
	}
	S.nq[m] = nq, S.nqt[m] = nqt, S.nqo[m] = nq - nqt;
}

// emission: one thread per quad of the group (j = rank in the group's quad list)
__device__ void emit_quads(const MeshParams &p, Small &S, const u64 *keys, const u32 *geo, int g0, int g1, int tid, int lane) {
	const u32 nquads = S.n_quads;
	for (u32 j0 = (u32)(tid - lane); j0 < nquads; j0 += kThreads) { // warp-uniform trip count
		const u32 j = j0 + lane;
		const bool valid = j < nquads;
		u32 s = 0, k = 0;
		u32 mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, mx[3] = {0, 0, 0};
		if (valid) {
			int lo = g0, hi = g1 - 1; // last mask with qpre[m] <= j
			while (lo < hi) {
				const int mid = (lo + hi + 1) >> 1;
				if (S.qpre[mid] <= j)
					lo = mid;
				else
					hi = mid - 1;
			}
			const int m = lo;
			const u32 i = j - S.qpre[m];
			const u32 slot = S.fcpre[m] - S.fcpre[g0] + i;
			const u64 key = keys[slot];
			const u32 g = geo[slot];
			const int axis = m / 66, rem = m - axis * 66, xa = rem >> 1, dir = rem & 1;
			const int ua = axis == 1 ? 2 : 1, va = axis == 0 ? 2 : 0; // BlockAlgo.hpp:70-81 with swizzle YZX
			const bool uv_order = dir ^ (axis != 2 ? 1 : 0);          // BlockMeshAlgo.hpp:382: quad_face_inv ^ ((U+1)%3 == V)
			const bool tr = (g >> 20) & 1;
			const u32 rank = g >> 21;
			s = tr ? 1u : 0u;
			const u32 q = (tr ? S.baset[m] : S.baseo[m]) + rank;
			const u32 other_before = (tr ? S.baseo[m] : S.baset[m]) + (i - rank);
			k = q / kQuadsPerMesh;
			const u32 u = g & 31u, v = (g >> 5) & 31u, width = ((g >> 10) & 31u) + 1, height = ((g >> 15) & 31u) + 1;
			// origin = ToXYZ<Axis>(xa,u,v); the two edge vectors (BlockMeshAlgo.hpp:380-390)
			const u32 P0[3] = {axis == 0 ? (u32)xa : v, axis == 1 ? (u32)xa : u, axis == 2 ? (u32)xa : (axis == 0 ? v : u)};
			const int a1 = uv_order ? ua : va, a2 = uv_order ? va : ua;
			const u32 l1 = uv_order ? height : width, l2 = uv_order ? width : height;
			const u32 uvec[3] = {a1 == 0 ? l1 : 0u, a1 == 1 ? l1 : 0u, a1 == 2 ? l1 : 0u};
			const u32 vvec[3] = {a2 == 0 ? l2 : 0u, a2 == 1 ? l2 : 0u, a2 == 2 ? l2 : 0u};
			const u32 code = (u32)(key >> 56);
			const u32 t16 = S.tab.tex16[code];
			const u32 face = (u32)(axis << 1) | (u32)dir;
			const u32 w1base = ((t16 & 0x1fffu) - 1u) | ((t16 >> 13) << 10) | (face << 13); // BlockVertex.hpp:16-17
			u32 sun[4], ao[4];
			uint2 vert[4];
#pragma unroll
			for (int c = 0; c < 4; ++c) {
				const u32 f = (u32)(key >> (14 * c)) & 0x3fffu;
				sun[c] = f & 63u, ao[c] = f >> 12;
				const u32 torch = (f >> 6) & 63u;
				u32 X = P0[0], Y = P0[1], Z = P0[2];
				if (c == 1 || c == 2)
					X += uvec[0], Y += uvec[1], Z += uvec[2];
				if (c == 2 || c == 3)
					X += vvec[0], Y += vvec[1], Z += vvec[2];
				vert[c].x = (X << 4) | (Y << 14) | (Z << 24) | ((u32)axis << 30);
				vert[c].y = w1base | (ao[c] << 16) | (sun[c] << 18) | (torch << 24);
			}
			uint4 *dst = reinterpret_cast<uint4 *>(p.vtx + S.voff + (s ? 4ull * S.Qo : 0ull) + 4ull * q);
			dst[0] = make_uint4(vert[0].x, vert[0].y, vert[1].x, vert[1].y);
			dst[1] = make_uint4(vert[2].x, vert[2].y, vert[3].x, vert[3].y);
			// Light4::GetFlip, BlockMeshAlgo.hpp:38-43
			const int d02 = abs((int)((ao[0] + 1) * (sun[0] + 1)) - (int)((ao[2] + 1) * (sun[2] + 1)));
			const int d13 = abs((int)((ao[1] + 1) * (sun[1] + 1)) - (int)((ao[3] + 1) * (sun[3] + 1)));
			store_indices(p, S.ioff, (s ? S.Qo : 0u) + q, q, d02 < d13);
#pragma unroll
			for (int c = 0; c < 3; ++c)
				mn[c] = P0[c] << 4, mx[c] = (P0[c] + uvec[c] + vvec[c]) << 4; // BlockMeshAlgo.hpp:404-409
			if (q == k * kQuadsPerMesh && k > 0)
				S.split_other[s][k - 1] = other_before; // who was ahead when this stream's mesh was cut
		}
		// AABB: reduce within lanes that target the same (stream, mesh) before touching shared memory
		const u32 grp = __match_any_sync(FULL, valid ? (s * 64u + k) : 0xffffffffu);
		u32 rmn[3], rmx[3];
#pragma unroll
		for (int c = 0; c < 3; ++c)
			rmn[c] = __reduce_min_sync(grp, mn[c]), rmx[c] = __reduce_max_sync(grp, mx[c]);
		if (valid && lane == __ffs(grp) - 1)
			aabb_atomic(S, s, k, rmn, rmx);
	}
}

// custom + dynamic faces of one in-chunk voxel (generate_custom_mesh, BlockMeshAlgo.hpp:238-303)
__device__ void emit_custom_voxel(const MeshParams &p, Small &S, const uint8_t *vox, const uint8_t *light, const DynRec *dyn, int x, int y, int z, u32 qo,
                                  u32 qt) {
	const u32 pr = vox[vox_index(x, y, z)];
	const u32 mesh = S.tab.pmesh[pr] - 1u;
	const hcm_custom_mesh *cm = p.meshes + mesh;
	const DevMeshInfo mi = p.mesh_info[mesh];
	DynRec rec{};
	u32 nfaces = cm->face_count, dynfaces = 0;
	if (mi.dynamic_kind == HCM_DYNAMIC_LIQUID_FLOW) {
		rec = dyn_load(dyn, x, y, z);
		dynfaces = rec.faces;
		nfaces = __popc(dynfaces);
	}
	const u32 base[3] = {(u32)x << 4, (u32)y << 4, (u32)z << 4};
	int cur_lf = -1;
	u32 w_ax = 0, u_ax = 0, v_ax = 0;
	L4 lo{}, hi{};
	for (u32 i = 0; i < nfaces; ++i) {
		CFace f;
		bool tr;
		if (mi.dynamic_kind == HCM_DYNAMIC_LIQUID_FLOW) {
			const int which = __ffs(dynfaces) - 1;
			dynfaces &= dynfaces - 1;
			flow_face(rec, S.tab.tex16[rec.flow_code], which, f);
			tr = rec.flow_trans;
		} else {
			const hcm_mesh_face *mf = cm->faces + i;
			f.axis = mf->axis, f.light_face = mf->light_face, f.render_face = mf->render_face, f.tex16 = mf->texture;
			const uint2 va_ = *reinterpret_cast<const uint2 *>(&mf->verts[0][0]), vb_ = *reinterpret_cast<const uint2 *>(&mf->verts[2][0]);
			const u32 vw[4] = {va_.x, va_.y, vb_.x, vb_.y};
#pragma unroll
			for (int c = 0; c < 4; ++c)
				f.v[c][0] = vw[c] & 0xff, f.v[c][1] = (vw[c] >> 8) & 0xff, f.v[c][2] = (vw[c] >> 16) & 0xff, f.v[c][3] = vw[c] >> 24;
			tr = (mi.trans_mask >> i) & 1;
		}
		if ((int)f.light_face != cur_lf) { // BlockMeshAlgo.hpp:247-268
			cur_lf = (int)f.light_face;
			w_ax = cur_lf >> 1, u_ax = (w_ax + 1) % 3, v_ax = (w_ax + 2) % 3;
			if (cur_lf & 1) {
				const u32 t = u_ax;
				u_ax = v_ax, v_ax = t;
			}
			const int step = 1 - ((cur_lf & 1) << 1);
			const int sx = w_ax == 0 ? step : 0, sy = w_ax == 1 ? step : 0, sz = w_ax == 2 ? step : 0;
			lo = light4_init(S, vox, light, cur_lf, x - sx, y - sy, z - sz);
			if (S.tab.pflags[vox[vox_index(x + sx, y + sy, z + sz)]] & PF_INDIRECT)
				hi = light4_init(S, vox, light, cur_lf, x, y, z);
			else
				hi = lo;
		}
		const u32 s = tr ? 1u : 0u;
		const u32 q = tr ? qt : qo;
		const u32 k = q / kQuadsPerMesh;
		if (q == k * kQuadsPerMesh && k > 0)
			S.split_other[s][k - 1] = tr ? qo : qt;
		u32 mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, mx[3] = {0, 0, 0};
		uint2 vert[4];
#pragma unroll
		for (int c = 0; c < 4; ++c) {
			const u32 px = base[0] + f.v[c][0], py = base[1] + f.v[c][1], pz = base[2] + f.v[c][2];
			mn[0] = min(mn[0], px), mn[1] = min(mn[1], py), mn[2] = min(mn[2], pz);
			mx[0] = max(mx[0], px), mx[1] = max(mx[1], py), mx[2] = max(mx[2], pz);
			auto comp = [&](u32 a) -> u32 { return a == 0 ? f.v[c][0] : (a == 1 ? f.v[c][1] : f.v[c][2]); };
			const u32 du = comp(u_ax), dv = comp(v_ax);
			u32 dw = min(comp(w_ax), 16u);
			if (cur_lf & 1)
				dw = (u32)(uint8_t)(16 - (int)dw);
			u32 ao = f.v[c][3];
			if (ao == 4)
				ao = (uint8_t)tri_interp(lo.ao, hi.ao, du, dv, dw);
			const u32 sun = (uint8_t)tri_interp(lo.sun, hi.sun, du, dv, dw);
			const u32 torch = (uint8_t)tri_interp(lo.torch, hi.torch, du, dv, dw);
			vert[c].x = px | (py << 10) | (pz << 20) | (f.axis << 30);
			vert[c].y = ((f.tex16 & 0x1fffu) - 1u) | ((f.tex16 >> 13) << 10) | (f.render_face << 13) | (ao << 16) | (sun << 18) | (torch << 24);
		}
		uint4 *dst = reinterpret_cast<uint4 *>(p.vtx + S.voff + (s ? 4ull * S.Qo : 0ull) + 4ull * q);
		dst[0] = make_uint4(vert[0].x, vert[0].y, vert[1].x, vert[1].y);
		dst[1] = make_uint4(vert[2].x, vert[2].y, vert[3].x, vert[3].y);
		store_indices(p, S.ioff, (s ? S.Qo : 0u) + q, q, true);
		aabb_atomic(S, s, k, mn, mx);
		if (tr)
			++qt;
		else
			++qo;
	}
}

// The flood columns of each thread depend on nothing but the thread index: they are computed once per context (the
// divisions by 34, 62 and 28 of slot_column cost ~600 instructions per warp and chunk when done in the mesh kernel).
__global__ void flood_slots_kernel(uint4 *out) {
	const int tid = threadIdx.x;
	u32 v[8] = {0xffffu, 0xffffu, 0xffffu, 0xffffu, 0xffffu, 0xffffu, 0xffffu, 0xffffu};
	static_assert(kColsPerThread <= 8, "flood slot table holds eight columns per thread");
	for (int k = 0; k < kColsPerThread; ++k) {
		int x = 0, z = 0; // domain coords: world x = x - 15
		bool ok = slot_column(tid, k, x, z);
		const int dx = x < 14 ? 14 - x : (x > 47 ? x - 47 : 0), dz = z < 14 ? 14 - z : (z > 47 ? z - 47 : 0);
		ok = ok && dx + dz <= 14; // farther columns can never reach the inner square
		if (ok)
			v[k] = (u32)((z + 1) * kSB + (x + 1)) | ((u32)(14 - dx - dz) << 12);
	}
	out[tid] = make_uint4(v[0] | (v[1] << 16), v[2] | (v[3] << 16), v[4] | (v[5] << 16), v[6] | (v[7] << 16));
}
void launch_flood_slots(uint4 *out, cudaStream_t st) { HCM_LAUNCH(flood_slots_kernel, 1, kThreads, 0, st, out); }

// The chunk's totals are known: claim its vertex, index and descriptor ranges.  One thread issues the three atomics back
// to back and only then looks at the results, so their round trips to the L2 overlap (three lanes with one atomic each
// measured slower: the divergent paths wait one after the other).
__device__ __forceinline__ void claim_output(const MeshParams &p, Small &S, int tid) {
	if (tid == 0) {
		S.Qo = S.run_o, S.Qt = S.run_t;
		const u64 nv = 4ull * (S.Qo + S.Qt), ni = 6ull * (S.Qo + S.Qt);
		const u32 nm = (S.Qo + kQuadsPerMesh - 1) / kQuadsPerMesh + (S.Qt + kQuadsPerMesh - 1) / kQuadsPerMesh;
		const u64 voff = atomicAdd(&p.heads->vtx, nv), ioff = atomicAdd(&p.heads->idx, ni);
		const u32 moff = atomicAdd(&p.heads->meshes, nm);
		S.voff = voff, S.ioff = ioff, S.moff = moff;
		if (voff + nv > p.vtx_cap || ioff + ni > p.idx_cap || moff + nm > p.mesh_cap)
			S.overflow = 1, atomicExch(&p.heads->overflow, 1u);
	}
}

template <int kVariant>
__global__ void __launch_bounds__(kThreads, 1) mesh_kernel(const __grid_constant__ MeshMaps maps, const MeshParams p) {
	HCM_DYNAMIC_SMEM(smem);
	uint8_t *vox = smem + SM_VOX;
	uint8_t *light = smem + SM_LIGHT;
	u32 *F = reinterpret_cast<u32 *>(smem + SM_F);
	u64 *Cpl = reinterpret_cast<u64 *>(smem + SM_CPLANE);
	uint8_t *region = smem + SM_REGION;
	Small &S = *reinterpret_cast<Small *>(smem + SM_SMALL);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const u32 lt = lanemask_lt();

	for (int i = tid; i < (int)(sizeof(DevTables) / 4); i += kThreads)
		reinterpret_cast<u32 *>(&S.tab)[i] = reinterpret_cast<const u32 *>(p.tables)[i];
	if (tid == 0) {
		mbar_init(&S.bar, 1);
		mbar_init(&S.bar_flood, 1);
	}
	__syncthreads();
	u32 phase = 0;
#ifdef HCM_PHASE_TIMING
	long long t_phase = clock64();
#endif
	DynRec *dyn = p.dyn_scratch + (size_t)blockIdx.x * kLightBytes;
	const u32 n_work = p.heads->n_work;

	// the two TMA boxes with the flood inputs of a task (thread 0)
	auto fetch_flood_inputs = [&](const TaskIn &t) {
		fence_proxy_async(); // generic accesses to these bytes come first
		mbar_expect_tx(&S.bar_flood, kFloodPassBytes + kFloodHeightBytes);
		tma_load_3d(smem + SM_FPASS, &maps.pass, t.cx * 32 - 16, t.cz * 32 - 15, t.cy - 1, &S.bar_flood);
		tma_load_3d(smem + SM_FHEIGHT, &maps.height, t.cx * 32 - 16, t.cz * 32 - 15, t.cy - 1, &S.bar_flood);
	};
	// The work cursor runs two items ahead: S.work_item is the current task, S.work_next the one after it, whose coordinates
	// thread 0 loads during the current task so that it can fetch that task's flood inputs before the current one has ended.
	if (tid == 0) {
		const u32 w0 = atomicAdd(&p.heads->cursor, 1u);
		S.work_item = w0, S.work_next = atomicAdd(&p.heads->cursor, 1u);
		if (w0 < n_work) {
			S.next_item = p.worklist[w0];
			S.next_task = p.tasks[S.next_item & 0x7fffffffu];
			fetch_flood_inputs(S.next_task);
		}
	}
	for (;;) {
		__syncthreads(); // previous task fully retired (smem reuse); S.work_item holds this CTA's next work item
		const u32 wi = S.work_item;
		PHASE_MARK(0); // task fetch
		if (wi >= n_work)
			break;
		// thread 0 claims the work item two after this one now and loads the next one's coordinates; the round trips hide behind
		// the task, and the cursor values are published behind the next barrier, when everybody has read the current one
		u32 wi_next = 0, wi_next2 = 0, item_next = 0;
		TaskIn ti_next{};
		if (tid == 0) {
			wi_next = S.work_next;
			wi_next2 = atomicAdd(&p.heads->cursor, 1u);
			if (wi_next < n_work) {
				item_next = p.worklist[wi_next];
				ti_next = p.tasks[item_next & 0x7fffffffu];
			}
		}
		const u32 item = S.next_item; // task index | torch sources among the 27 chunks << 31 (classify_kernel)
		const u32 task = item & 0x7fffffffu;
		const bool any_torch = item >> 31;
		const TaskIn ti = S.next_task;

		// ---- phase 0: halo gather -- one TMA box of the dense property grid ---------------------------------------
		// (the flood's inputs -- pass words and sunlight heights of the 62 x 62 columns x 3 chunk layers around the task -- are
		// two more boxes, already on their way: fetch_flood_inputs)
		if (tid == 0) {
			fence_proxy_async();
			mbar_expect_tx(&S.bar, kVoxBytes);
			tma_load_3d(vox, &maps.vox, ti.cx * 32, ti.cz * 32 - 2, ti.cy * 32 - 1, &S.bar); // stored column = x + kXMargin
		}
		for (int i = tid; i < 2 * kMaxSplits * 6; i += kThreads)
			(&S.aabb[0][0][0])[i] = (i % 6) < 3 ? 0xffffffffu : 0u; // dep/AABB.hpp:11
		if (tid == 0) {
			S.n_custom = 0, S.any_dyn = 0, S.overflow = 0, S.n_cand = 0, S.cand_overflow = 0;
		}
		PHASE_MARK(1); // task set-up
		// ---- light first: the flood reads the y-packed planes in global memory, not the voxel box, so it runs while the TMA
		// box of this task is still in flight (a task without any face pays for a flood it does not need: classify_kernel
		// leaves next to none of those) --------------------------------------------------------------------------------
		const bool want_light = p.debug_light != nullptr;
#ifdef HCM_PHASE_TIMING
		const long long tq0_ = clock64();
#endif
		__syncthreads(); // the per-task fields of S are reset, S.work_item has been read by everybody
#ifdef HCM_PHASE_TIMING
		const long long tq1_ = clock64();
#endif
		if (tid == 0)
			S.work_item = wi_next, S.work_next = wi_next2, S.next_task = ti_next, S.next_item = item_next; // loaded a task set-up ago: no long-lived registers
		{
			u64 *sbuf = reinterpret_cast<u64 *>(smem + SM_SUNBUF);
			for (int i = tid; i < 2 * kSB * kSB; i += kThreads)
				sbuf[i] = 0; // both buffers: zero border and interior
#ifdef HCM_PHASE_TIMING
			const long long tq2_ = clock64();
#endif
			ColInfo ci;
			ci.X = (u32)p.w.X;
			ci.g0 = (u32)(((size_t)(ti.cy - 1) * p.w.Z + (size_t)(ti.cz * 32 - 15)) * p.w.X + (size_t)(ti.cx * 32 - 15));
			u64 P[kColsPerThread], sun0[kColsPerThread]; // passable cells; sunlit cells (the sources of the sunlight flood)
			// The columns' pass words and heights were fetched by the two TMA boxes issued with the halo box; slots without a
			// column read box entry 0.
			const uint4 slots = __ldg(p.flood_slots + tid);
			u32 bidx[kColsPerThread];
#pragma unroll
			for (int k = 0; k < kColsPerThread; ++k) {
				const u32 v = ((k < 2 ? slots.x : (k < 4 ? slots.y : (k < 6 ? slots.z : slots.w))) >> (16 * (k & 1))) & 0xffffu;
				const bool ok = v != 0xffffu;
				ci.c[k] = ok ? (int)((v & 0xfffu) | ((v >> 12) << 16)) : -1;
				bidx[k] = ok ? (u32)(ci.idx(k) - kSB) : 0u; // sbuf index (z+1)*64 + (x+1) -> box index z*64 + (x+1)
			}
			static_assert(kSB == kFloodBoxX, "the flood buffers and the TMA boxes share their row pitch");
			const u32 *fpass = reinterpret_cast<const u32 *>(smem + SM_FPASS);
			const int8_t *fheight = reinterpret_cast<const int8_t *>(smem + SM_FHEIGHT);
			mbar_wait(&S.bar_flood, phase);
			u32 pw[kColsPerThread][3];
			int hw[kColsPerThread][3];
#pragma unroll
			for (int k = 0; k < kColsPerThread; ++k)
#pragma unroll
				for (int d = 0; d < 3; ++d) {
					pw[k][d] = fpass[bidx[k] + d * kFloodBoxCols];
					hw[k][d] = fheight[bidx[k] + d * kFloodBoxCols];
				}
#pragma unroll
			for (int k = 0; k < kColsPerThread; ++k) {
				u32 sl[3];
#pragma unroll
				for (int d = 0; d < 3; ++d) {
					const int h = hw[k][d];
					sl[d] = __funnelshift_lc(0u, 0xffffffffu, (u32)max(h, 0)); // sunlit: y >= height (all for h <= 0, none for h >= 32), Chunk.hpp:73-76
				}
				const bool ok = ci.c[k] >= 0;
				P[k] = ok ? combine3(pw[k][0], pw[k][1], pw[k][2]) : 0ull;
				sun0[k] = ok ? combine3(sl[0], sl[1], sl[2]) : 0ull;
			}
#ifdef HCM_PHASE_TIMING
			const long long tq3_ = clock64();
#endif
			__syncthreads();
#ifdef HCM_PHASE_TIMING
			if (tid == 64) {
				PHASE_COUNT(61, tq1_ - tq0_);        // wait at the set-up barrier (thread 0's flag loads)
				PHASE_COUNT(62, tq2_ - tq1_);        // zeroing
				PHASE_COUNT(63, tq3_ - tq2_);        // plane + height loads
				PHASE_COUNT(60, clock64() - tq3_);   // wait at the barrier after the loads
			}
#endif
			// Step 0 of the sunlight flood reads its sources from the buffer it does not write (its own entry only).  The
			// word is what the column holds after step 0 anyway, so the double-buffer protocol is not disturbed.
#pragma unroll
			for (int k = 0; k < kColsPerThread; ++k)
				if (ci.c[k] >= 0)
					sbuf[kSB * kSB + ci.idx(k)] = sun0[k];
			PHASE_MARK(6); // flood setup (plane loads)
			flood_channel<false>(p, ci, P, sbuf, light, tid);
			if (any_torch && p.w.n_levels > 0) {
				u64 *tbuf = reinterpret_cast<u64 *>(region + RG_SBUF); // the staged inputs are in registers by now
				for (int i = tid; i < 2 * kSB * kSB; i += kThreads)
					tbuf[i] = 0;
				__syncthreads();
				flood_channel<true>(p, ci, P, tbuf, light, tid);
			}
			if (want_light) // canonical order (x fastest, then z, then y) for hcm_debug_light
				for (int i = tid; i < kLightBytes; i += kThreads) {
					const int x = i % kLS - 1, z = (i / kLS) % kLS - 1, y = i / (kLS * kLS) - 1;
					p.debug_light[(size_t)task * kLightBytes + i] = light[lvol_index(x, y, z)];
				}
		}

		PHASE_MARK(7); // flood + unpack
		mbar_wait(&S.bar, phase);
		phase ^= 1;
		PHASE_MARK(17); // TMA wait
		for (int i = tid; i < 66 * kFS; i += kThreads)
			F[i] = 0; // axis-0 masks are filled with atomicOr; the others are overwritten (the flood inputs were parked here until now)

		// ---- phase 1: class bit planes (bit x+1 of row (y,z)); one lane builds one row from three 16-byte loads -------
		{
			u64 *Opl = reinterpret_cast<u64 *>(region + RG_O), *Epl = reinterpret_cast<u64 *>(region + RG_E), *Dpl = reinterpret_cast<u64 *>(region + RG_D);
			u32 ncust = 0, anyd = 0;
			for (int row = tid; row < kInnerCols; row += kThreads) {
				const int yy = row / kInner, zz = row % kInner; // y + 1, z + 1
				const uint4 *r = reinterpret_cast<const uint4 *>(vox + (yy * kVZ + zz + 1) * kVX);
				const uint4 q0 = r[0], q1 = r[1], q2 = r[2];
				const u32 w[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
				// Eight voxels at a time: the four class bits of a voxel are spread to bit 0 of the four bytes of a word
				// (spread4) and shifted to the voxel's place, so that byte k of the group's word is class k of its 8 voxels --
				// four instructions per voxel instead of two per voxel AND class; byte permutes assemble the row words.
				u32 acc[5] = {0, 0, 0, 0, 0};
#pragma unroll
				for (int i = 0; i < 34; ++i) { // stored byte i + 3 <-> world x = i - 1 <-> plane bit i
					const u32 pr = (w[(i + 3) >> 2] >> (8 * ((i + 3) & 3))) & 0xffu;
					const u32 cls = ((u32)S.tab.pflags[pr] >> 2) & 15u; // bit0 O, bit1 E, bit2 C, bit3 D
					acc[i >> 3] += spread4(cls) << (i & 7);
				}
				u32 lo[4], hi[4];
#pragma unroll
				for (int k = 0; k < 4; ++k) {
					const u32 pick = (u32)(k | ((4 + k) << 4)); // byte k of the first operand, byte k of the second
					lo[k] = __byte_perm(__byte_perm(acc[0], acc[1], pick), __byte_perm(acc[2], acc[3], pick), 0x5410u);
					hi[k] = (acc[4] >> (8 * k)) & 0xffu;
				}
				Opl[row] = ((u64)hi[0] << 32) | lo[0], Epl[row] = ((u64)hi[1] << 32) | lo[1];
				Cpl[row] = ((u64)hi[2] << 32) | lo[2], Dpl[row] = ((u64)hi[3] << 32) | lo[3];
				if (yy >= 1 && yy <= 32 && zz >= 1 && zz <= 32)
					ncust += __popc((lo[2] >> 1) | (hi[2] << 31));
				anyd |= lo[3] | hi[3];
			}
			ncust = __reduce_add_sync(FULL, ncust);
			anyd = __reduce_or_sync(FULL, anyd);
			if (lane == 0 && (ncust | anyd)) {
				atomicAdd(&S.n_custom, ncust);
				if (anyd)
					S.any_dyn = 1;
			}
		}
		__syncthreads();

		PHASE_MARK(2); // planes
		// ---- phase 1b: liquid-flow voxels of [-1,32]^3 -> heights + per-voxel face textures --------------------------
		if (S.any_dyn) {
			const u64 *Dpl = reinterpret_cast<const u64 *>(region + RG_D);
			for (int row = tid; row < kInnerCols; row += kThreads) {
				u64 w = Dpl[row];
				const int y = row / kInner - 1, z = row % kInner - 1;
				while (w) {
					const int x = __ffsll((long long)w) - 2;
					w &= w - 1;
					const DynRec r = flow_record(p, S, vox, ti, x, y, z);
					*reinterpret_cast<u64 *>(dyn + cell_index(x, y, z)) = *reinterpret_cast<const u64 *>(&r);
				}
			}
			__syncthreads();
		}

		PHASE_MARK(3); // liquid
		// ---- phase 2: faces from the planes -------------------------------------------------------------------------
		// A pair (a,b) of voxels across a slice: O|E (opaque cube next to textureless voxel) always shows a's face and
		// never b's (Show(): a non-empty, b = kNone is transparent, a opaque); O|O and E|E never show anything; only
		// pairs with a voxel that is neither (water, leaves, glass, cactus, flowing water ...) need the exact Show().
		// Those pairs are first collected into a list so that the exact test runs with every lane busy.
		{
			const u64 *Opl = reinterpret_cast<const u64 *>(region + RG_O), *Epl = reinterpret_cast<const u64 *>(region + RG_E);
			u32 *clist = reinterpret_cast<u32 *>(region + RG_CLIST);
			const uint8_t *tf = S.tab.texflags;
			// exact Show() of one pair; entry = axis<<16 | slice<<10 | u<<5 | v
			auto exact = [&](u32 ent) {
				const int axis = ent >> 16, xa = (ent >> 10) & 63, u = (ent >> 5) & 31, v = ent & 31;
				const int ax = axis == 0 ? xa - 1 : v, ay = axis == 1 ? xa - 1 : u, az = axis == 2 ? xa - 1 : (axis == 0 ? v : u);
				const int bx = axis == 0 ? xa : v, by = axis == 1 ? xa : u, bz = axis == 2 ? xa : (axis == 0 ? v : u);
				// two voxels of the same property whose six textures are one texture (water, leaves, glass ...): Show() sees
				// a == b on both sides and shows nothing -- most candidate pairs of an ocean or a tree crown end here
				const u32 pa = vox[vox_index(ax, ay, az)];
				if (pa == vox[vox_index(bx, by, bz)] && (S.tab.pmisc[pa] & 4u))
					return;
				const u32 ta = tex_code(S, vox, dyn, ax, ay, az, axis * 2), tb = tex_code(S, vox, dyn, bx, by, bz, axis * 2 + 1);
				if (xa != 0 && tex_show(tf, ta, tb))
					atomicOr(&F[mask_id(axis, xa, 0) * kFS + u], 1u << v);
				if (xa != 32 && tex_show(tf, tb, ta))
					atomicOr(&F[mask_id(axis, xa, 1) * kFS + u], 1u << v);
			};
			for (int unit = tid; unit < 1024 + 2 * 1056; unit += kThreads) {
				if (unit < 1024) { // axis 0: row (y,z); bit xa of the shifted words <-> voxel pair (xa-1, xa)
					const int y = unit >> 5, z = unit & 31;
					const u64 o = Opl[(y + 1) * kInner + z + 1], e = Epl[(y + 1) * kInner + z + 1];
					const u64 x_ = ~(o | e) & ((1ull << 34) - 1);
					u64 s0 = o & (e >> 1) & 0x1fffffffeull; // + faces: slices 1..32
					u64 s1 = e & (o >> 1) & 0x0ffffffffull; // - faces: slices 0..31
					u64 cand = (x_ | (x_ >> 1)) & ((1ull << 33) - 1);
					while (s0) {
						const int xa = __ffsll((long long)s0) - 1;
						s0 &= s0 - 1;
						atomicOr(&F[mask_id(0, xa, 0) * kFS + y], 1u << z);
					}
					while (s1) {
						const int xa = __ffsll((long long)s1) - 1;
						s1 &= s1 - 1;
						atomicOr(&F[mask_id(0, xa, 1) * kFS + y], 1u << z);
					}
					if (cand) {
						u32 pos = atomicAdd(&S.n_cand, (u32)__popcll(cand));
						while (cand) {
							const u32 xa = __ffsll((long long)cand) - 1;
							cand &= cand - 1;
							const u32 ent = (xa << 10) | ((u32)y << 5) | (u32)z;
							if (pos < (u32)kCandCap)
								clist[pos] = ent;
							else
								exact(ent); // list full (stress chunks): evaluate in place
							++pos;
						}
					}
				} else {
					const int axis = unit < 1024 + 1056 ? 1 : 2;
					const int k = unit - (axis == 1 ? 1024 : 1024 + 1056);
					const int xa = k >> 5, u = k & 31;
					// axis 1: slice y = xa, row u = z ; axis 2: slice z = xa, row u = y ; v = x for both
					const int ya = axis == 1 ? xa - 1 : u, za = axis == 1 ? u : xa - 1;
					const int yb = axis == 1 ? xa : u, zb = axis == 1 ? u : xa;
					const int ra = (ya + 1) * kInner + za + 1, rb = (yb + 1) * kInner + zb + 1;
					const u32 oa = (u32)(Opl[ra] >> 1), ea = (u32)(Epl[ra] >> 1), ob = (u32)(Opl[rb] >> 1), eb = (u32)(Epl[rb] >> 1);
					// every (mask,row) word of axes 1 and 2 is initialised by exactly one unit
					F[mask_id(axis, xa, 0) * kFS + u] = xa != 0 ? (oa & eb) : 0u;
					F[mask_id(axis, xa, 1) * kFS + u] = xa != 32 ? (ea & ob) : 0u;
					u32 cand = ~(oa | ea) | ~(ob | eb);
					if (cand) {
						u32 pos = atomicAdd(&S.n_cand, (u32)__popc(cand));
						while (cand) {
							const u32 x = __ffs(cand) - 1;
							cand &= cand - 1;
							const u32 ent = ((u32)axis << 16) | ((u32)xa << 10) | ((u32)u << 5) | x;
							if (pos < (u32)kCandCap)
								clist[pos] = ent;
							else
								S.cand_overflow = 1; // handled below, after F is initialised everywhere
							++pos;
						}
					}
				}
			}
			__syncthreads();
			const u32 nc = min(S.n_cand, (u32)kCandCap);
			for (u32 i = tid; i < nc; i += kThreads)
				exact(clist[i]);
			if (S.cand_overflow) {
				// stress chunks: more exact pairs on axes 1/2 than the list holds.  Redo those units in place (every pair again:
				// atomicOr is idempotent), now that all F words have been initialised.
				for (int unit = 1024 + tid; unit < 1024 + 2 * 1056; unit += kThreads) {
					const int axis = unit < 1024 + 1056 ? 1 : 2;
					const int k = unit - (axis == 1 ? 1024 : 1024 + 1056);
					const int xa = k >> 5, u = k & 31;
					const int ya = axis == 1 ? xa - 1 : u, za = axis == 1 ? u : xa - 1;
					const int yb = axis == 1 ? xa : u, zb = axis == 1 ? u : xa;
					const int ra = (ya + 1) * kInner + za + 1, rb = (yb + 1) * kInner + zb + 1;
					u32 cand = ~((u32)(Opl[ra] >> 1) | (u32)(Epl[ra] >> 1)) | ~((u32)(Opl[rb] >> 1) | (u32)(Epl[rb] >> 1));
					while (cand) {
						const u32 x = __ffs(cand) - 1;
						cand &= cand - 1;
						exact(((u32)axis << 16) | ((u32)xa << 10) | ((u32)u << 5) | x);
					}
				}
			}
		}
		__syncthreads();

		PHASE_MARK(4); // faces
		// ---- phase 3: faces per mask, prefix; early out ----------------------------------------------------------------
		for (int m = warp; m < kMasks; m += kWarps) {
			const u32 c = __reduce_add_sync(FULL, (u32)__popc(F[m * kFS + lane]));
			if (lane == 0)
				S.fc[m] = c;
		}
		__syncthreads();
		if (warp == 0) {
			u32 carry = 0;
			for (int b = 0; b < kMasks; b += 32) {
				const u32 v = (b + lane < kMasks) ? S.fc[b + lane] : 0u;
				const u32 ex = warp_excl_scan(v, lane);
				if (b + lane < kMasks)
					S.fcpre[b + lane] = carry + ex;
				carry += __shfl_sync(FULL, ex + v, 31);
			}
			if (lane == 0)
				S.fcpre[kMasks] = carry;
		}
		__syncthreads();
		PHASE_MARK(5); // face counts + prefix
		const u32 n_faces = S.fcpre[kMasks];
		const u32 n_custom = S.n_custom;
		if (n_faces == 0 && n_custom == 0 && !want_light) {
			if (tid == 0) {
				p.task_out[task] = hcm_task_result{0u, 0u};
				if (S.work_item < n_work)
					fetch_flood_inputs(S.next_task); // the phase region was last read before the barrier above
			}
			continue;
		}

		// ---- phase 5: custom faces per row and stream (count), exclusive prefix in (y,z) row order -------------------
		if (n_custom) {
			constexpr int kRPT = 1024 / kThreads; // (y,z) rows per thread
			static_assert(kRPT == 1 || kRPT == 2, "phase 5 handles 1 or 2 rows per thread");
			u32 co2[2] = {0, 0}, ct2[2] = {0, 0};
#pragma unroll
			for (int j = 0; j < kRPT; ++j) {
				const int r = kRPT * tid + j, y = r >> 5, z = r & 31;
				u32 w = (u32)(Cpl[(y + 1) * kInner + z + 1] >> 1);
				while (w) {
					const int x = __ffs(w) - 1;
					w &= w - 1;
					const DevMeshInfo mi = p.mesh_info[S.tab.pmesh[vox[vox_index(x, y, z)]] - 1];
					if (mi.dynamic_kind == HCM_DYNAMIC_LIQUID_FLOW) {
						const DynRec rec = dyn_load(dyn, x, y, z);
						const u32 nf = __popc((u32)rec.faces);
						if (rec.flow_trans)
							ct2[j] += nf;
						else
							co2[j] += nf;
					} else
						co2[j] += mi.n_opaque, ct2[j] += mi.n_trans;
				}
			}
			const u32 so = co2[0] + co2[1], st = ct2[0] + ct2[1];
			const u32 exo = warp_excl_scan(so, lane), ext = warp_excl_scan(st, lane);
			if (lane == 31)
				S.wtmp[0][warp] = exo + so, S.wtmp[1][warp] = ext + st;
			__syncthreads();
			if (warp == 0) {
				const u32 vo = lane < kWarps ? S.wtmp[0][lane] : 0u, vt = lane < kWarps ? S.wtmp[1][lane] : 0u;
				const u32 eo = warp_excl_scan(vo, lane), et = warp_excl_scan(vt, lane);
				if (lane < kWarps)
					S.wtmp[0][lane] = eo, S.wtmp[1][lane] = et;
				if (lane == kWarps - 1)
					S.Co = eo + vo, S.Ct = et + vt;
			}
			__syncthreads();
			const u32 bo = S.wtmp[0][warp] + exo, bt = S.wtmp[1][warp] + ext;
			S.crow_o[kRPT * tid] = bo, S.crow_t[kRPT * tid] = bt;
			if (kRPT == 2)
				S.crow_o[2 * tid + 1] = bo + co2[0], S.crow_t[2 * tid + 1] = bt + ct2[0];
		} else if (tid == 0)
			S.Co = 0, S.Ct = 0;
		__syncthreads();

		PHASE_MARK(8); // custom count
		// ---- phases 6/7: keys -> merge -> (count) -> allocate -> emit ---------------------------------------------------
		u64 *keys = reinterpret_cast<u64 *>(region + RG_KEYS);
		u32 *geo = reinterpret_cast<u32 *>(region + RG_GEO);
		const bool single = n_faces <= (u32)kGroupCap;
		const int n_pass = single ? 1 : 2;
		for (int pass = 0; pass < n_pass; ++pass) {
			const bool emit = pass == n_pass - 1;
			if (tid == 0)
				S.run_o = S.Co, S.run_t = S.Ct;
			__syncthreads();
			int g0 = 0;
			while (g0 < kMasks) {
				// largest g1 with faces(g0..g1) <= cap, at least one mask
				int g1;
				{
					const u32 lim = S.fcpre[g0] + (u32)kGroupCap;
					int lo = g0 + 1, hi = kMasks;
					while (lo < hi) { // upper bound on the monotone prefix
						const int mid = (lo + hi + 1) >> 1;
						if (S.fcpre[mid] <= lim)
							lo = mid;
						else
							hi = mid - 1;
					}
					g1 = lo;
				}
				const u32 nf = S.fcpre[g1] - S.fcpre[g0];
				if (nf) {
					// (a) face locations into the key slots, in mask/row/column order
					for (int m = g0 + warp; m < g1; m += kWarps) {
						if (!S.fc[m])
							continue;
						u32 fw = F[m * kFS + lane];
						const u32 rows = __ballot_sync(FULL, fw != 0);
						if (lane == 0)
							S.rowmask[m] = rows;
						u32 at = S.fcpre[m] - S.fcpre[g0] + warp_excl_scan(__popc(fw), lane);
						while (fw) {
							const u32 v = __ffs(fw) - 1;
							fw &= fw - 1;
							keys[at++] = ((u64)m << 10) | ((u64)lane << 5) | v;
						}
					}
					__syncthreads();
					PHASE_MARK(9); // fill
					// (b) one thread per visible face: texture + Light4 -> key
					for (u32 f = tid; f < nf; f += kThreads) {
						const u32 loc = (u32)keys[f];
						const int m = loc >> 10, u = (loc >> 5) & 31, v = loc & 31;
						const int axis = m / 66, rem = m - axis * 66, xa = rem >> 1, dir = rem & 1;
						const int sl = xa - (dir ? 0 : 1); // the voxel owning the face: xa-1 for the + face, xa for the - face
						const int q[3] = {axis == 0 ? sl : v, axis == 1 ? sl : u, axis == 2 ? sl : (axis == 0 ? v : u)};
						const int face = axis * 2 + dir;
						const u32 code = tex_code(S, vox, dyn, q[0], q[1], q[2], face);
						const L4 l4 = light4_init(S, vox, light, face, q[0], q[1], q[2]);
						keys[f] = pack_key(l4, code);
					}
					__syncthreads();
					PHASE_MARK(10); // keys
					// (c) greedy merge, one warp per mask.  A mask costs ~ its non-empty rows, and the heavy ones (terrain tops)
					// cluster on every other mask id, so warps pull masks from a shared counter, heavy masks first.
					if (tid == 0)
						S.n_order[0] = 0, S.n_order[1] = 0, S.n_order[2] = 0, S.next_mask = 0;
					__syncthreads();
					if (g0 + tid < g1) {
						const int m = g0 + tid;
						const u32 c = S.fc[m];
						if (c == 0)
							S.nq[m] = 0, S.nqo[m] = 0, S.nqt[m] = 0;
						else {
							const int cls = c >= 128 ? 0 : (c > kThreadMergeMax ? 1 : 2);
							S.order[cls][atomicAdd(&S.n_order[cls], 1u)] = (uint8_t)m;
						}
					}
					__syncthreads();
					{
						// tiny masks: one thread each, on the last warps; everything else: warps pull from a shared counter
						const u32 n_tiny = S.n_order[2];
						const int tiny_warps = min((int)((n_tiny + kTinyPerWarp - 1) / kTinyPerWarp), kWarps);
#ifdef HCM_PHASE_TIMING
						const long long tw0_ = clock64();
#endif
						if (warp >= kWarps - tiny_warps) {
							// warp w of the tiny group takes masks w, w + tiny_warps, ... on its first lanes
							const int tw = warp - (kWarps - tiny_warps);
							for (u32 j = (u32)(tw + lane * tiny_warps); j < n_tiny && lane < 32; j += 32u * tiny_warps)
								merge_mask_thread(S, F, keys, geo, S.order[2][j], g0);
							__syncwarp();
#ifdef HCM_PHASE_TIMING
							if (lane == 0) {
								PHASE_COUNT(35, clock64() - tw0_); // tiny-mask warps
								PHASE_COUNT(36, 1);
							}
#endif
						}
						const u32 nh = S.n_order[0], ntot = nh + S.n_order[1];
						for (;;) {
							u32 j = 0;
							if (lane == 0)
								j = atomicAdd(&S.next_mask, 1u);
							j = __shfl_sync(FULL, j, 0);
							if (j >= ntot)
								break;
							const int m = j < nh ? S.order[0][j] : S.order[1][j - nh];
#ifdef HCM_PHASE_TIMING
							const long long tm0_ = clock64();
							const u32 nr_ = __popc(__ballot_sync(FULL, F[m * kFS + lane] != 0));
#endif
							merge_mask<kVariant>(S, F, keys, geo, m, g0, lane);
#ifdef HCM_PHASE_TIMING
							if (lane == 0) {
								const unsigned long long d_ = (unsigned long long)(clock64() - tm0_);
								PHASE_COUNT(24, d_);
								PHASE_COUNT(22, nr_);
								PHASE_COUNT(20, 1);
								atomicMax(&g_phase_cycles[33], d_);
							}
#endif
						}
#ifdef HCM_PHASE_TIMING
						if (lane == 0)
							PHASE_COUNT(34, clock64() - tw0_); // per-warp time inside the merge section, summed over warps
#endif
					}
					__syncthreads();
					PHASE_MARK(11); // merge
					// (d) ordered offsets: exclusive scans of the per-mask quad counts (both streams, and all quads)
					{
						const int m = g0 + tid;
						const u32 vo = m < g1 ? S.nqo[m] : 0u, vt = m < g1 ? S.nqt[m] : 0u;
						const u32 eo = warp_excl_scan(vo, lane), et = warp_excl_scan(vt, lane);
						if (lane == 31 && warp < 8)
							S.wtmp[0][warp] = eo + vo, S.wtmp[1][warp] = et + vt;
						__syncthreads();
						u32 co = 0, ct = 0;
						for (int w2 = 0; w2 < warp && w2 < 8; ++w2)
							co += S.wtmp[0][w2], ct += S.wtmp[1][w2];
						if (m < g1) {
							S.baseo[m] = S.run_o + co + eo, S.baset[m] = S.run_t + ct + et;
							S.qpre[m] = co + eo + ct + et;
						}
						__syncthreads();
						if (tid == 0) {
							u32 to = 0, tt = 0;
							for (int w2 = 0; w2 < 8; ++w2)
								to += S.wtmp[0][w2], tt += S.wtmp[1][w2];
							S.run_o += to, S.run_t += tt, S.n_quads = to + tt;
							PHASE_COUNT(23, to + tt); // quads
						}
					}
					__syncthreads();
				}
				if (single) {
					// totals are known after the only group: allocate now
					claim_output(p, S, tid);
					__syncthreads();
				}
				PHASE_MARK(13); // alloc
				if (emit && nf && !S.overflow)
					emit_quads(p, S, keys, geo, g0, g1, tid, lane);
				__syncthreads();
				g0 = g1;
			}
			if (!single && pass == 0) {
				claim_output(p, S, tid);
				__syncthreads();
			}
		}

		if (tid == 0 && S.work_item < n_work)
			fetch_flood_inputs(S.next_task); // the phase region is idle from here on (every pass ends in a barrier)
		PHASE_MARK(14); // emit
		// ---- custom / dynamic faces: warp per (y,z) row, lane per voxel --------------------------------------------------
		if (n_custom && !S.overflow) {
			for (int r = warp; r < 1024; r += kWarps) {
				const int y = r >> 5, z = r & 31;
				const u32 w = (u32)(Cpl[(y + 1) * kInner + z + 1] >> 1);
				if (!w)
					continue;
				u32 no = 0, nt = 0;
				const bool mine = (w >> lane) & 1;
				if (mine) {
					const DevMeshInfo mi = p.mesh_info[S.tab.pmesh[vox[vox_index(lane, y, z)]] - 1];
					if (mi.dynamic_kind == HCM_DYNAMIC_LIQUID_FLOW) {
						const DynRec rec = dyn_load(dyn, lane, y, z);
						const u32 nf = __popc((u32)rec.faces);
						if (rec.flow_trans)
							nt = nf;
						else
							no = nf;
					} else
						no = mi.n_opaque, nt = mi.n_trans;
				}
				const u32 qo = S.crow_o[r] + warp_excl_scan(no, lane), qt = S.crow_t[r] + warp_excl_scan(nt, lane);
				if (mine)
					emit_custom_voxel(p, S, vox, light, dyn, lane, y, z, qo, qt);
			}
		}
		__syncthreads();

		PHASE_MARK(15); // custom emit
		// ---- mesh descriptors in the order BlockMeshAlgo::Generate returns them (BlockMeshAlgo.hpp:397-403, :515-518) ----
		if (tid == 0) {
			if (S.overflow) {
				p.task_out[task] = hcm_task_result{0u, 0u};
			} else {
				const u32 Q[2] = {S.Qo, S.Qt};
				const u32 M[2] = {(Q[0] + kQuadsPerMesh - 1) / kQuadsPerMesh, (Q[1] + kQuadsPerMesh - 1) / kQuadsPerMesh};
				const u32 full[2] = {M[0] ? M[0] - 1 : 0u, M[1] ? M[1] - 1 : 0u};
				u32 out = S.moff;
				auto put = [&](u32 s, u32 k) {
					hcm_mesh_desc d;
					const u32 nq = min(Q[s] - k * kQuadsPerMesh, kQuadsPerMesh);
					d.task = task, d.transparent = s;
					d.vertex_offset = S.voff + (s ? 4ull * Q[0] : 0ull) + 4ull * kQuadsPerMesh * k;
					d.index_offset = S.ioff + 6ull * ((s ? Q[0] : 0u) + (u64)kQuadsPerMesh * k);
					d.vertex_count = 4 * nq, d.index_count = 6 * nq;
					for (int c = 0; c < 3; ++c)
						d.aabb_min[c] = S.aabb[s][k][c], d.aabb_max[c] = S.aabb[s][k][3 + c];
					p.mesh_out[out++] = d;
				};
				u32 i = 0, j = 0; // cut meshes, interleaved by the time of the cut
				while (i < full[0] || j < full[1]) {
					bool take_opaque;
					if (i >= full[0])
						take_opaque = false;
					else if (j >= full[1])
						take_opaque = true;
					else // the opaque quad that triggers cut i is quad number 16383*(i+1); it precedes transparent trigger j iff
					     // more opaque quads than that were already emitted when trigger j arrived
						take_opaque = S.split_other[1][j] > kQuadsPerMesh * (i + 1);
					if (take_opaque)
						put(0, i++);
					else
						put(1, j++);
				}
				if (M[0])
					put(0, M[0] - 1);
				if (M[1])
					put(1, M[1] - 1);
				p.task_out[task] = hcm_task_result{S.moff, M[0] + M[1]};
			}
		}
		PHASE_MARK(16); // descriptors
	}
}

// ================================================================================================================
// device-side sink: every placed mesh is copied from the batch arenas to the range the caller allocated for
// it in its own vertex / index buffer -- what MeshHandle's two staging buffers and vkCmdCopyBuffer do in the reference
// (client/include/client/mesh/MeshHandle.hpp:75-87, MeshCluster.hpp:76-118).  One CTA per mesh (grid-stride), 16-byte copies
// where source and destination allow it: arena offsets are multiples of 4 vertices / 6 indices, the caller's offsets are
// whatever its allocator returned (alignment sizeof(Vertex) = 8 and sizeof(Index) = 2, MeshCluster.hpp:44-55).
// ================================================================================================================
__global__ void __launch_bounds__(256) scatter_kernel(const hcm_mesh_placement *__restrict__ placements, u32 n, const hcm_mesh_desc *__restrict__ meshes,
                                                      const u64 *__restrict__ vtx, const uint16_t *__restrict__ idx, u64 *__restrict__ vdst,
                                                      uint16_t *__restrict__ idst) {
	for (u32 i = blockIdx.x; i < n; i += gridDim.x) {
		const hcm_mesh_placement pl = placements[i];
		const u64 voff = meshes[pl.mesh].vertex_offset, ioff = meshes[pl.mesh].index_offset;
		const u32 nv = meshes[pl.mesh].vertex_count, ni = meshes[pl.mesh].index_count;
		const u64 *vs = vtx + voff;
		u64 *vd = vdst + pl.vertex_offset;
		if ((((size_t)vs | (size_t)vd) & 15) == 0) { // nv is a multiple of 4
			const uint4 *s4 = reinterpret_cast<const uint4 *>(vs);
			uint4 *d4 = reinterpret_cast<uint4 *>(vd);
			for (u32 k = threadIdx.x; k < nv / 2; k += blockDim.x)
				d4[k] = s4[k];
		} else
			for (u32 k = threadIdx.x; k < nv; k += blockDim.x)
				vd[k] = vs[k];
		const uint16_t *is = idx + ioff;
		uint16_t *id = idst + pl.index_offset;
		if ((((size_t)is | (size_t)id) & 3) == 0) { // ni is a multiple of 6
			const u32 *s2 = reinterpret_cast<const u32 *>(is);
			u32 *d2 = reinterpret_cast<u32 *>(id);
			for (u32 k = threadIdx.x; k < ni / 2; k += blockDim.x)
				d2[k] = s2[k];
		} else
			for (u32 k = threadIdx.x; k < ni; k += blockDim.x)
				id[k] = is[k];
	}
}

// ================================================================================================================
// on-device qualification of the kernel variants (hcm.h, "Kernel variants"): position-salted hashes of what a launch wrote.
// The sum of mix64(value ^ salt(position)) does not depend on the order in which threads add their terms, and the mesh hash
// does not depend on where in the arenas a task's meshes were placed (the allocation order differs from run to run).
// ================================================================================================================
__device__ __forceinline__ u64 mix64(u64 x) {
	x ^= x >> 33, x *= 0xff51afd7ed558ccdull;
	x ^= x >> 33, x *= 0xc4ceb9fe1a85ec53ull;
	return x ^ (x >> 33);
}
__device__ __forceinline__ u64 warp_sum64(u64 h) {
#pragma unroll
	for (int d = 16; d; d >>= 1)
		h += __shfl_sync(FULL, h, (int)((threadIdx.x & 31) ^ d));
	return h;
}

// one CTA per upload entry, the derive kernel's own thread mapping: everything derive_kernel writes for the chunk
__global__ void __launch_bounds__(256) derive_digest_kernel(DevWorld w, const UploadEntry *__restrict__ entries, const int8_t *__restrict__ raw_heights,
                                                            const u64 *__restrict__ expect, u64 *__restrict__ out, u32 *__restrict__ mismatches) {
	__shared__ u64 s_sum;
	const int tid = threadIdx.x;
	const UploadEntry e = entries[blockIdx.x];
	if (tid == 0)
		s_sum = 0;
	__syncthreads();
	const int z = tid >> 3, x4 = (tid & 7) << 2;
	const size_t gx = (size_t)e.cx * 32 + x4, gz = (size_t)e.cz * 32 + z;
	const uint8_t *prow = w.prop8 + (((size_t)e.cy * 32) * w.Z + gz) * w.XP + gx + kXMargin;
	const size_t pstride = (size_t)w.Z * w.XP;
	const size_t col = ((size_t)e.cy * w.Z + gz) * w.X + gx;
	const size_t plane = (size_t)w.wy * w.Z * w.X;
	u64 h = 0;
	u32 salt = (u32)tid << 8;
	for (int y = 0; y < 32; ++y)
		h += mix64(*reinterpret_cast<const u32 *>(prow + y * pstride) ^ ((u64)(salt + y) << 32));
	const uint4 pw = *reinterpret_cast<const uint4 *>(w.pass + col);
	h += mix64(pw.x ^ ((u64)(salt + 32) << 32)) + mix64(pw.y ^ ((u64)(salt + 33) << 32)) + mix64(pw.z ^ ((u64)(salt + 34) << 32)) + mix64(pw.w ^ ((u64)(salt + 35) << 32));
	for (int l = 0; l < w.n_levels; ++l) {
		const uint4 tw = *reinterpret_cast<const uint4 *>(w.torch + l * plane + col);
		h += mix64(tw.x ^ ((u64)(salt + 40 + 4 * l) << 32)) + mix64(tw.y ^ ((u64)(salt + 41 + 4 * l) << 32)) + mix64(tw.z ^ ((u64)(salt + 42 + 4 * l) << 32)) +
		     mix64(tw.w ^ ((u64)(salt + 43 + 4 * l) << 32));
	}
	h += mix64(*reinterpret_cast<const u32 *>(w.height + col) ^ ((u64)(salt + 60) << 32));
	// the raw slot as well: a host-uniform entry has it written by the derive kernel (first and last 16 bytes per thread row suffice
	// to see a wrong fill value; the full 64 KB are the upload's, not the kernel's, for every other entry)
	if (e.flags & UE_UNIFORM) {
		const uint16_t *raw = w.raw + (size_t)e.slot * HCM_CHUNK_VOXELS;
		for (int y = 0; y < 32; ++y) {
			const uint2 r = *reinterpret_cast<const uint2 *>(raw + x4 + (y * 32 + z) * 32);
			h += mix64(((u64)r.y << 32 | r.x) ^ (u64)(salt + 64 + y));
		}
	}
	if (tid == 0) {
		const size_t ci = ((size_t)e.cy * w.wz + e.cz) * w.wx + e.cx;
		h += mix64(w.cflags[ci] ^ (0x5151ull << 32)) + mix64((u32)w.slot[ci] ^ (0x5252ull << 32));
	}
	(void)raw_heights;
	h = warp_sum64(h);
	if ((tid & 31) == 0)
		atomicAdd(reinterpret_cast<unsigned long long *>(&s_sum), (unsigned long long)h);
	__syncthreads();
	if (tid == 0) {
		out[blockIdx.x] = s_sum;
		if (expect && expect[blockIdx.x] != s_sum)
			atomicAdd(mismatches, 1u);
	}
}

// one warp per task: mesh count, then per mesh (in the task's order) flags, counts, AABB, vertex bytes, index bytes -- not the offsets
__global__ void __launch_bounds__(256) mesh_digest_kernel(const hcm_task_result *__restrict__ task_out, const hcm_mesh_desc *__restrict__ meshes,
                                                          const u64 *__restrict__ vtx, const uint16_t *__restrict__ idx, u32 n, u32 mesh_cap,
                                                          unsigned long long vtx_cap, unsigned long long idx_cap, const u64 *__restrict__ expect,
                                                          u64 *__restrict__ out, u32 *__restrict__ mismatches) {
	const u32 t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (t >= n)
		return;
	const hcm_task_result tr = task_out[t];
	u64 h = lane == 0 ? mix64((u64)tr.n_meshes + 0x9e3779b97f4a7c15ull) : 0ull;
	bool bad = tr.n_meshes > 2u * kMaxSplits || (u64)tr.first_mesh + tr.n_meshes > mesh_cap; // a descriptor range no kernel of this library writes
	for (u32 k = 0; k < tr.n_meshes && !bad; ++k) {
		const hcm_mesh_desc d = meshes[tr.first_mesh + k];
		if (d.vertex_count > HCM_MESH_MAX_VERTICES || d.vertex_offset + d.vertex_count > vtx_cap || d.index_offset + d.index_count > idx_cap ||
		    d.index_count > 6u * (HCM_MESH_MAX_VERTICES / 4)) {
			bad = true;
			break;
		}
		if (lane == 0) {
			h += mix64(((u64)k << 40) ^ ((u64)d.transparent << 32) ^ d.vertex_count) + mix64(((u64)k << 40) ^ (1ull << 36) ^ d.index_count) +
			     mix64(((u64)k << 40) ^ (2ull << 36) ^ d.task);
			for (int c = 0; c < 3; ++c)
				h += mix64(((u64)k << 40) ^ ((u64)(3 + c) << 36) ^ d.aabb_min[c]) + mix64(((u64)k << 40) ^ ((u64)(6 + c) << 36) ^ d.aabb_max[c]);
		}
		const u64 *v = vtx + d.vertex_offset;
		for (u32 i = lane; i < d.vertex_count; i += 32)
			h += mix64(v[i] ^ mix64(((u64)k << 32) | i));
		const uint16_t *ix = idx + d.index_offset;
		for (u32 i = lane; i < d.index_count; i += 32)
			h += mix64((u64)ix[i] ^ mix64(((u64)(k + 64) << 32) | i));
	}
	h = warp_sum64(h);
	if (bad)
		h = ~0ull; // and it will not equal the other run's hash unless that one is broken in the same way: counted below
	if (lane == 0) {
		out[t] = h;
		if (expect && (expect[t] != h || bad))
			atomicAdd(mismatches, 1u);
	}
}

void launch_derive_digest(const DevWorld &w, const UploadEntry *entries, u32 n, const int8_t *raw_heights, const u64 *expect, u64 *out, u32 *mismatches,
                          cudaStream_t st) {
	if (n)
		HCM_LAUNCH(derive_digest_kernel, n, 256, 0, st, w, entries, raw_heights, expect, out, mismatches);
}
void launch_mesh_digest(const hcm_task_result *task_out, const hcm_mesh_desc *meshes, const u64 *vtx, const uint16_t *idx, u32 n, u32 mesh_cap,
                        unsigned long long vtx_cap, unsigned long long idx_cap, const u64 *expect, u64 *out, u32 *mismatches, cudaStream_t st) {
	if (n)
		HCM_LAUNCH(mesh_digest_kernel, (n + 7) / 8, 256, 0, st, task_out, meshes, vtx, idx, n, mesh_cap, vtx_cap, idx_cap, expect, out, mismatches);
}

// ================================================================================================================
// host-callable launchers (used by hcm_api.cu)
// ================================================================================================================
void launch_scatter(const hcm_mesh_placement *placements, u32 n, const hcm_mesh_desc *meshes, const u64 *vtx, const uint16_t *idx, void *vdst, void *idst,
                    int grid, cudaStream_t st) {
	if (n)
		HCM_LAUNCH(scatter_kernel, min((u32)grid, n), 256, 0, st, placements, n, meshes, vtx, idx, (u64 *)vdst, (uint16_t *)idst);
}
void launch_derive(const DevWorld &w, const UploadEntry *entries, u32 n, const uint8_t *prop_index, const DevTables *tables, const int8_t *raw_heights,
                   int variant, bool trial, cudaStream_t st) {
	if (!n)
		return;
	// (the emulation's launch macro takes the kernel as ONE macro argument: no comma inside the template argument list)
#define HCM_DERIVE_CASE(form, kernel) \
	case form: HCM_LAUNCH(kernel, n, 256, HCM_CHUNK_VOXELS * 2, st, w, entries, prop_index, tables, raw_heights); break
	if (trial) {
		constexpr auto t1 = derive_kernel<1, true>, t2 = derive_kernel<2, true>, t3 = derive_kernel<3, true>;
		switch (variant) {
			HCM_DERIVE_CASE(1, t1);
			HCM_DERIVE_CASE(2, t2);
			HCM_DERIVE_CASE(3, t3);
		default: HCM_LAUNCH(derive_kernel<0>, n, 256, HCM_CHUNK_VOXELS * 2, st, w, entries, prop_index, tables, raw_heights); break;
		}
	} else {
		switch (variant) {
			HCM_DERIVE_CASE(1, derive_kernel<1>);
			HCM_DERIVE_CASE(2, derive_kernel<2>);
			HCM_DERIVE_CASE(3, derive_kernel<3>);
		default: HCM_LAUNCH(derive_kernel<0>, n, 256, HCM_CHUNK_VOXELS * 2, st, w, entries, prop_index, tables, raw_heights); break;
		}
	}
#undef HCM_DERIVE_CASE
}
int derive_variant_count() { return kDeriveForms - 1; }
int mesh_variant_count() { return 2; }
void launch_clear(const DevWorld &w, const UploadEntry *entries, u32 n, cudaStream_t st) {
	if (n)
		HCM_LAUNCH(clear_kernel, n, 256, 0, st, w, entries);
}
void launch_edit(const DevWorld &w, const hcm_block_edit *edits, u32 n, const uint2 *columns, u32 n_columns, u32 *result, const uint8_t *prop_index,
                 const DevTables *tables, int8_t *raw_heights, hcm_sunlight_change *changes, u32 *n_changes, u32 cap, cudaStream_t st) {
	if (!n)
		return;
	HCM_LAUNCH(edit_blocks_kernel, (n + 127) / 128, 128, 0, st, w, edits, n, result);
	HCM_LAUNCH(sunlight_chain_kernel, (n_columns + 127) / 128, 128, 0, st, w, edits, columns, n_columns, result, prop_index, tables, raw_heights, changes, n_changes, cap);
}
void launch_classify(const DevWorld &w, const TaskIn *tasks, u32 n, u32 *worklist, BatchHeads *heads, hcm_task_result *task_out, int force_all,
                     cudaStream_t st) {
	if (n)
		HCM_LAUNCH(classify_kernel, (n + 255) / 256, 256, 0, st, w, tasks, n, worklist, heads, task_out, force_all);
}
#ifdef HCM_PHASE_TIMING
extern "C" __attribute__((visibility("default"))) int hcm_debug_phase_cycles(unsigned long long *out, int reset) {
	cudaMemcpyFromSymbol(out, g_phase_cycles, sizeof(g_phase_cycles));
	if (reset) {
		unsigned long long z[64] = {};
		cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z));
	}
	return 0;
}
#endif
cudaError_t mesh_kernel_prepare() {
	cudaError_t e = cudaFuncSetAttribute(derive_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, HCM_CHUNK_VOXELS * 2);
	if (e == cudaSuccess)
		e = cudaFuncSetAttribute(derive_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, HCM_CHUNK_VOXELS * 2);
	if (e == cudaSuccess)
		e = cudaFuncSetAttribute(derive_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, HCM_CHUNK_VOXELS * 2);
	if (e == cudaSuccess)
		e = cudaFuncSetAttribute(derive_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, HCM_CHUNK_VOXELS * 2);
	if (e == cudaSuccess)
		e = cudaFuncSetAttribute(derive_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, HCM_CHUNK_VOXELS * 2);
	if (e == cudaSuccess)
		e = cudaFuncSetAttribute(derive_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, HCM_CHUNK_VOXELS * 2);
	if (e == cudaSuccess)
		e = cudaFuncSetAttribute(derive_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, HCM_CHUNK_VOXELS * 2);
	if (e == cudaSuccess)
		e = cudaFuncSetAttribute(mesh_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
	if (e == cudaSuccess)
		e = cudaFuncSetAttribute(mesh_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
	if (e == cudaSuccess)
		e = cudaFuncSetAttribute(mesh_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
	return e;
}
int mesh_kernel_smem_bytes() { return kSmemBytes; }
void launch_mesh(const MeshMaps &maps, const MeshParams &p, int grid, int variant, cudaStream_t st) {
	if (variant == 1)
		HCM_LAUNCH(mesh_kernel<1>, grid, kThreads, kSmemBytes, st, maps, p);
	else if (variant == 2)
		HCM_LAUNCH(mesh_kernel<2>, grid, kThreads, kSmemBytes, st, maps, p);
	else
		HCM_LAUNCH(mesh_kernel<0>, grid, kThreads, kSmemBytes, st, maps, p);
}

} // namespace hcm
