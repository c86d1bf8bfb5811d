// Device-side data model of the B200 chunk mesher.  See DESIGN.md for the HBM layout and the kernel pipeline.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hcm.h"

// Kernel launch and dynamic shared memory go through two macros so that the CPU test tier (tests/cusim: an emulation of the
// CUDA execution model that compiles these very sources with g++) can substitute its own; the CUDA build is what ships.
#ifndef HCM_LAUNCH
#define HCM_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif
#ifndef HCM_DYNAMIC_SMEM
#define HCM_DYNAMIC_SMEM(name) extern __shared__ __align__(128) uint8_t name[]
#endif

namespace hcm {

// ---- launch geometry ------------------------------------------------------------------------------------------
#ifndef HCM_MESH_THREADS
#define HCM_MESH_THREADS 1024
#endif
constexpr int kThreads = HCM_MESH_THREADS; // mesh kernel: one persistent CTA per SM (shared memory bound); the phases are
                                           // latency bound, so as many warps as a CTA can hold
constexpr int kWarps = kThreads / 32;

// ---- shared-memory voxel volume (property index per voxel), filled by ONE 3-D TMA box -----------------------------
// z in [-2,33] (the +-2 reach is only read by the liquid-flow generator), y in [-1,32].  TMA needs the innermost box
// extent AND the innermost start coordinate to be multiples of 16 bytes (measured: a start of 32c-2 raises "illegal
// instruction"), so the dense grid stores world column x at x + kXMargin and the box starts at 32c: x in [-4,43].
constexpr int kVX = 48, kVZ = 36, kVY = 34;
constexpr int kXMargin = 4;
constexpr int kVoxBytes = kVX * kVZ * kVY; // 58752
__host__ __device__ constexpr int vox_index(int x, int y, int z) { return ((y + 1) * kVZ + (z + 2)) * kVX + (x + kXMargin); }

// ---- canonical cell order on [-1,32]^3 (debug light output, liquid scratch records) ---------------------------------
constexpr int kLS = 34;
constexpr int kLightBytes = kLS * kLS * kLS; // 39304
__host__ __device__ constexpr int cell_index(int x, int y, int z) { return (x + 1) + ((z + 1) + (y + 1) * kLS) * kLS; }

// ---- shared-memory light volume on [-1,32]^3: byte = torch << 4 | sun (block/include/block/Light.hpp:20-24) -------
// y is the fastest axis (padded to 36) so that a flood column unpacks into nine aligned 32-bit stores.
constexpr int kLY = 36;
constexpr int kLvolBytes = kLY * kLS * kLS; // 41616
__host__ __device__ constexpr int lvol_index(int x, int y, int z) { return (y + 1) + kLY * ((x + 1) + kLS * (z + 1)); }

// ---- face masks: one 32x32 bit mask per (axis, slice 0..32, direction) --------------------------------------------
constexpr int kMasks = 3 * 33 * 2; // 198
__host__ __device__ constexpr int mask_id(int axis, int xa, int dir) { return (axis * 33 + xa) * 2 + dir; }

// ---- greedy-merge working set: faces of a group of consecutive masks ----------------------------------------------
constexpr int kGroupCap = 6144; // faces per group (key 8 B + geometry 4 B each)

// ---- flood domain: 62 x 62 columns (x,z in [-15,46]) of 62-bit words (bit = y + 15) -------------------------------
constexpr int kFD = 62;
constexpr int kFCols = kFD * kFD;       // 3844
constexpr int kInnerSlots = (1156 + kThreads - 1) / kThreads; // inner columns per thread (their light planes live in registers)
constexpr int kOuterSlots = (2688 + kThreads - 1) / kThreads; // outer columns per thread
constexpr int kColsPerThread = kInnerSlots + kOuterSlots;
constexpr int kInner = 34;              // columns x,z in [-1,32] carry the result
constexpr int kInnerCols = kInner * kInner;
// The flood's inputs (pass words, sunlight heights) of a task arrive as two TMA boxes of 64 x 62 columns x 3 chunk layers:
// box column 0 is world x = 32 cx - 16 (16-byte aligned for both element sizes), domain column x sits at box column x + 1.
constexpr int kFloodBoxX = 64;
constexpr int kFloodBoxCols = kFloodBoxX * kFD;              // 3968 columns per chunk layer
constexpr int kFloodPassBytes = kFloodBoxCols * 3 * 4;       // 47616
constexpr int kFloodHeightBytes = kFloodBoxCols * 3;         // 11904

struct MeshMaps {
	CUtensorMap vox, pass, height;
};

constexpr uint32_t kQuadsPerMesh = HCM_MESH_MAX_VERTICES / 4; // 16383 quads fill a BlockMesh (BlockMeshAlgo.hpp:397-403)
constexpr int kMaxSplits = 40;                                 // >= ceil(622592 / 16383) meshes per stream

// ---- per-property flags (DevTables::pflags) ------------------------------------------------------------------------
constexpr uint8_t PF_INDIRECT = 1;  // Block::GetIndirectLightPass  (transparency != opaque), Block.hpp:70
constexpr uint8_t PF_DIRECT = 2;    // GetVerticalLightPass || collision != solid, BlockMeshAlgo.hpp:119,124
constexpr uint8_t PF_SOLID = 4;     // class O: six non-empty, non-transparent, non-liquid face textures, no custom mesh
constexpr uint8_t PF_EMPTY = 8;     // class E: six empty face textures, no dynamic mesh
constexpr uint8_t PF_CUSTOM = 16;   // has a custom mesh
constexpr uint8_t PF_DYNAMIC = 32;  // custom mesh is dynamic (liquid flow)
constexpr uint8_t PF_LIQ_SOURCE = 64;
constexpr uint8_t PF_LIQ_FLOW = 128;

static_assert(PF_INDIRECT == 1 && PF_SOLID == 4 && PF_EMPTY == 8, "derive_kernel relies on these bit positions");

// texture code flags (DevTables::texflags): HCM_TEX_* plus
constexpr uint8_t TF_EMPTY = 8; // BlockTexture::Empty(): id == 0

struct DevTables {
	uint8_t ptex[256][8];  // property row -> texture code per face (codes index tex16/texflags; 0 = BlockTexture{} )
	uint8_t pflags[256];   // PF_*
	uint8_t pmesh[256];    // custom mesh index + 1, 0 = none
	uint8_t pmisc[256];    // bit0: liquid obstacle, bit1: vertical light pass (transparency == transparent), bit2: the six face
	                       // textures are one and the same and the mesh is not dynamic, bits 4..7: light level
	uint16_t tex16[256];   // code -> BlockTexture::GetData()
	uint8_t texflags[256]; // code -> HCM_TEX_* | TF_EMPTY
};
static_assert(sizeof(DevTables) == 3584, "DevTables layout");

// per custom mesh: static face counts per stream + liquid parameters, for the count pass
struct DevMeshInfo {
	uint8_t n_opaque, n_trans; // static faces per output stream
	uint8_t dynamic_kind;
	uint8_t flow_code;         // texture code of liquid_flow_tex
	uint8_t flow_trans;        // liquid_flow_tex uses the transparent pass
	uint8_t height_unit, variant, _pad;
	uint16_t trans_mask;       // bit i: static face i goes to the transparent stream
	uint16_t _pad2;
};
static_assert(sizeof(DevMeshInfo) == 12, "DevMeshInfo layout");

// chunk summary flags (DevWorld::cflags), written by the derive kernel
constexpr uint32_t CF_ALL_EMPTY = 1u << 0;   // every voxel class E
constexpr uint32_t CF_ALL_SOLID = 1u << 1;   // every voxel class O
constexpr uint32_t CF_HAS_CUSTOM = 1u << 2;
constexpr uint32_t CF_HAS_DYNAMIC = 1u << 3;
constexpr uint32_t CF_HAS_TORCH = 1u << 4;
constexpr uint32_t CF_RESIDENT = 1u << 5;
constexpr uint32_t CF_FACE_SOLID0 = 1u << 8; // + f: boundary layer on side f is all class O
constexpr uint32_t CF_FACE_EMPTY0 = 1u << 16;

constexpr int kMaxLightLevels = 4; // distinct non-zero BlockProperty::light_level values supported

struct DevWorld {
	int ox, oy, oz; // window origin (chunk coordinates)
	int wx, wy, wz; // window extent (chunks)
	int X, Z;       // voxel extents of the dense grids along x and z (32*wx, 32*wz)
	int XP;         // row pitch of prop8: 32*wx + 16 (world column x is stored at x + kXMargin)
	uint8_t *prop8;       // [32*wy][Z][XP]  property index per voxel (x fastest, then z, then y -- the chunk order)
	uint32_t *pass;       // [wy][Z][X]      bit y_local: voxel lets indirect light pass
	uint32_t *torch;      // [levels][wy][Z][X] bit y_local: voxel emits light level levels[l]
	int8_t *height;       // [wy][Z][X]      sunlight height of the column within its chunk
	uint32_t *cflags;     // [wy][wz][wx]
	int32_t *slot;        // [wy][wz][wx]    raw slot or -1
	const uint16_t *raw;  // [max_chunks][32768] uploaded Block data (needed for raw meta of flowing liquids)
	int n_levels;
	uint8_t levels[kMaxLightLevels]; // descending
};

struct UploadEntry {
	int32_t slot;
	int16_t cx, cy, cz; // window-relative chunk coordinates
	uint16_t value;     // UE_UNIFORM: the Block value every voxel of the chunk holds
	uint32_t flags;
};
constexpr uint32_t UE_UNIFORM = 1; // host found the chunk uniform: its 64 KB were not copied, the kernel fills the raw slot

struct TaskIn {
	int16_t cx, cy, cz; // window-relative chunk coordinates
	int16_t _pad;
};

// dyn record of a liquid-flow voxel (global scratch, per CTA): heights, per-face texture mask, emitted side faces
struct __align__(8) DynRec {
	uint8_t h[4];     // h0..h3, Liquid.hpp:204-207
	uint8_t texmask;  // bit f: face f carries the flow texture (else kNone), Liquid.hpp:211,287-292
	uint8_t faces;    // bit0 top, bit1 left, bit2 right, bit3 back, bit4 front (emission order of flow_mesh_gen)
	uint8_t flow_code;
	uint8_t flow_trans;
};

struct BatchHeads {
	unsigned long long vtx;   // vertices allocated
	unsigned long long idx;   // indices allocated
	unsigned int meshes;      // mesh descriptors allocated
	unsigned int overflow;    // set when an allocation did not fit
	unsigned int n_work;      // tasks appended to the work list by classify
	unsigned int cursor;      // next work item
};

struct MeshParams {
	DevWorld w;
	const DevTables *tables;
	const hcm_custom_mesh *meshes;
	const DevMeshInfo *mesh_info;
	const TaskIn *tasks;
	const uint32_t *worklist;
	BatchHeads *heads;
	hcm_task_result *task_out;
	hcm_mesh_desc *mesh_out;
	uint64_t *vtx;
	uint16_t *idx;
	unsigned long long vtx_cap, idx_cap;
	unsigned int mesh_cap;
	DynRec *dyn_scratch;   // [gridDim.x][34^3]
	const uint4 *flood_slots; // [kThreads] the flood columns of each thread, eight 16-bit words: sbuf index | last relevant step << 12, 0xffff = none
	uint8_t *debug_light;  // optional [n_tasks][39304]
};

} // namespace hcm
