/*
 * hcm.h -- C ABI of the B200 chunk mesher (libhcm.so).
 *
 * Drop-in boundary for HyperCraft's client mesh task.  A re-written
 * `ChunkTaskRunner<ChunkTaskType::kMesh>::Run` (reference client/src/ChunkMeshTask.cpp:41-127) or a batched
 * dispatcher draining many popped `ChunkTaskRunnerData<kMesh>` (client/include/client/ChunkMeshTask.inl:17-29)
 * calls these entry points instead of BlockLightAlgo::PropagateLight + BlockMeshAlgo::Generate, and feeds the
 * returned meshes to `WorldRenderer::PushChunkMesh` (client/include/client/WorldRenderer.hpp:69-87) unchanged.
 *
 * Plain pointers and sizes only: no C++ types, no torch types, no exceptions.  Every function returns
 * HCM_OK (0) or a negative hcm_status; hcm_last_error() gives the text.  One hcm_ctx per GPU; calls on one
 * ctx must be externally serialised (the dispatcher thread owns it) -- this replaces the reference's N worker
 * threads each owning a runner (client/include/client/ChunkTaskPool.hpp:187-194).
 *
 * There is NO CPU fallback: hcm_create fails with HCM_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef HCM_H
#define HCM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------------------
 * Constants fixed by the reference data model
 * ---------------------------------------------------------------------------------------------------------- */
#define HCM_CHUNK_SIZE 32            /* kChunkSize, common/include/common/Size.hpp:8 */
#define HCM_CHUNK_VOXELS 32768       /* Block m_blocks[32^3], client/include/client/Chunk.hpp:95 */
#define HCM_CHUNK_COLUMNS 1024       /* InnerPos1 m_sunlight_heights[32^2], Chunk.hpp:96 */
#define HCM_MAX_MESH_FACES 16        /* block::kBlockMeshMaxFaceCount, block/include/block/BlockMesh.hpp:37 */
#define HCM_MESH_MAX_VERTICES 65532u /* a BlockMesh is cut when cur_vertex + 4 > UINT16_MAX, BlockMeshAlgo.hpp:397-403 */

typedef enum hcm_status {
	HCM_OK = 0,
	HCM_ERR_INVALID = -1,     /* bad argument / malformed registry */
	HCM_ERR_CUDA = -2,        /* CUDA runtime/driver failure (text in hcm_last_error) */
	HCM_ERR_UNSUPPORTED = -3, /* registry uses a feature the device path cannot mirror (unknown dynamic mesh, ...) */
	HCM_ERR_NOT_RESIDENT = -4,/* a task's chunk or one of its 26 neighbours was never uploaded (reference: Pop returns
	                             std::nullopt, ChunkMeshTask.cpp:29-31) */
	HCM_ERR_OUT_OF_WINDOW = -5,/* chunk position outside the device-resident world window */
	HCM_ERR_CAPACITY = -6,    /* more resident chunks / tasks than configured */
	HCM_ERR_STATE = -7        /* call sequence error (fetch without batch, ...) */
} hcm_status;

/* ------------------------------------------------------------------------------------------------------------
 * Registry: flat mirror of the reference's compile-time tables
 *   block::Block::kBlockDataProperties[65536]  (block/include/block/Block.hpp:35-41, block/generator/Block.cpp.in:19-30)
 *   block::BlockProperty                        (block/include/block/BlockProperty.hpp:17-29)
 *   block::BlockMesh / BlockMeshFace            (block/include/block/BlockMesh.hpp:15-49)
 *   texture::BlockTexture flag tables           (texture/internal/generated/block_texture_transparency.inl:1-3)
 * A function pointer (BlockMesh::p_dynamic_mesh_func) cannot cross the ABI; the exporter names the generator
 * instead (dynamic_kind) and passes its template parameters.
 * ---------------------------------------------------------------------------------------------------------- */
#define HCM_TEX_TRANSPARENT 1u /* BlockTexture::IsTransparent, texture/include/texture/BlockTexture.hpp:33-36 */
#define HCM_TEX_LIQUID 2u      /* BlockTexture::IsLiquid, :37-40 */
#define HCM_TEX_TRANS_PASS 4u  /* BlockTexture::UseTransparentPass, :41-44 */

#define HCM_DYNAMIC_NONE 0u
#define HCM_DYNAMIC_LIQUID_FLOW 1u /* block::flow_mesh_gen<Config,Variant>, block/block/public/Liquid.hpp:171-282 */
#define HCM_OBSTACLE_ID_NOT_AIR 0u /* WaterTrait::IsObstacle: id != kAir, block/block/public/Water.hpp:13 */

#define HCM_LIQUID_NONE 0u
#define HCM_LIQUID_SOURCE 1u /* block id == Config::kSourceID */
#define HCM_LIQUID_FLOW 2u   /* block id == Config::kFlowID */

typedef struct hcm_block_prop {
	uint16_t tex[6];      /* BlockTexture::GetData() for faces Right,Left,Top,Bottom,Front,Back (BlockFace.hpp:10) */
	uint8_t transparency; /* BlockTransparency: 0 opaque, 1 semi, 2 transparent (BlockProperty.hpp:15) */
	uint8_t light_level;  /* LightLvl 0..15 */
	uint8_t collision;    /* BlockCollision: 0 none, 1 solid, 2 water (BlockProperty.hpp:16) */
	uint8_t custom_mesh;  /* 0 = regular cube, else index+1 into hcm_registry::meshes */
	uint8_t liquid_role;  /* HCM_LIQUID_*: what flow_mesh_gen's get_nei_height sees for this block id (Liquid.hpp:179-184) */
	uint8_t liquid_obstacle; /* Config::IsObstacle(block) (Water.hpp:13) */
	uint8_t _pad[2];
} hcm_block_prop; /* 20 bytes */

typedef struct hcm_mesh_face {
	uint8_t axis;        /* BlockMeshFace::axis */
	uint8_t light_face;  /* BlockMeshFace::light_face */
	uint8_t render_face; /* BlockMeshFace::render_face */
	uint8_t _pad0;
	uint16_t texture;    /* BlockTexture::GetData() */
	uint16_t _pad1;
	uint8_t verts[4][4]; /* x, y, z in 1/16 voxel, ao (4 = auto, BlockMesh.hpp:22) */
} hcm_mesh_face; /* 24 bytes */

typedef struct hcm_custom_mesh {
	uint32_t face_count;   /* static faces (0 for dynamic meshes, Liquid.hpp:285-288) */
	uint32_t dynamic_kind; /* HCM_DYNAMIC_* */
	uint8_t liquid_source_id, liquid_flow_id; /* Config::kSourceID / kFlowID (informational) */
	uint8_t liquid_height_unit;               /* 16 / Config::kFlowDistance (get_flow_height, Liquid.hpp:164-169) */
	uint8_t liquid_variant;                   /* template parameter Variant of flow_mesh_gen */
	uint16_t liquid_flow_tex;                 /* BlockTexture{Config::kFlowTex}.GetData() */
	uint8_t liquid_obstacle_rule;             /* HCM_OBSTACLE_* (informational; the per-block bit is in hcm_block_prop) */
	uint8_t _pad;
	hcm_mesh_face faces[HCM_MAX_MESH_FACES];
} hcm_custom_mesh; /* 400 bytes */

typedef struct hcm_registry {
	uint32_t n_props;    /* distinct property rows, <= 256 */
	uint32_t n_meshes;   /* distinct custom meshes, <= 255 */
	uint32_t n_textures; /* length of tex_flags (texture ids 0..n_textures-1; id 0 = kNone) */
	uint32_t _pad;
	const uint8_t *prop_index;     /* [65536]: Block::GetData() -> row of props (the flat table of Block.hpp:41) */
	const hcm_block_prop *props;   /* [n_props] */
	const hcm_custom_mesh *meshes; /* [n_meshes] */
	const uint8_t *tex_flags;      /* [n_textures] HCM_TEX_* bits */
} hcm_registry;

/* Serialised registry ("blob"): header, prop_index[65536], props[], meshes[], tex_flags[]. */
#define HCM_REGISTRY_MAGIC 0x4752434Du /* "MCRG" */
#define HCM_REGISTRY_VERSION 1u
typedef struct hcm_registry_blob_header {
	uint32_t magic, version, n_props, n_meshes, n_textures, _pad[3];
} hcm_registry_blob_header; /* 32 bytes */

/* Point `out` into a blob held by the caller (no copy). */
int hcm_registry_from_blob(const void *blob, size_t size, hcm_registry *out);

/* ------------------------------------------------------------------------------------------------------------
 * Context: one per GPU.  Holds the device-resident world window and the per-batch output arenas.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct hcm_ctx hcm_ctx;

typedef struct hcm_config {
	int32_t device;           /* CUDA ordinal */
	int16_t origin[3];        /* chunk position (ChunkPos3) of the window's minimum corner */
	uint16_t extent[3];       /* window size in chunks per axis; positions in [origin, origin+extent) are addressable */
	uint32_t max_chunks;      /* resident raw-chunk capacity (0 = extent product) */
	uint32_t max_tasks;       /* max tasks per hcm_mesh_batch (0 = extent product) */
	uint64_t vertex_capacity; /* initial output arena size in vertices (0 = default); grows on demand, never truncates */
} hcm_config;

int hcm_create(const hcm_registry *registry, const hcm_config *config, hcm_ctx **out);
void hcm_destroy(hcm_ctx *ctx);
const char *hcm_last_error(const hcm_ctx *ctx); /* ctx may be NULL: error of the last failed hcm_create */
const char *hcm_status_string(int status);

/* Pinned host memory for upload sources (so the host->device copies are truly asynchronous). */
void *hcm_host_alloc(size_t bytes);
void hcm_host_free(void *p);

/* ------------------------------------------------------------------------------------------------------------
 * Chunk residency.  Inputs are exactly what a reference Chunk exposes:
 *   blocks : Chunk::GetBlockData(), 32768 x uint16 (Block::GetData(): id | meta << 8), index x + (y*32 + z)*32
 *            (Chunk.hpp:50, common/include/common/Position.hpp:98-100)
 *   heights: Chunk::GetSunlightHeight(idx), 1024 x int8, index x + z*32; sunlit(x,y,z) = y >= height (Chunk.hpp:65-76)
 *   pos    : Chunk::GetPosition() (ChunkPos3 = i16vec3)
 * The call enqueues the host->device copy and the derivation kernel on the ctx stream and returns; the host
 * arrays must stay valid until hcm_sync / hcm_mesh_fetch.  Re-uploading a position replaces it (block edits).
 * Positions must be distinct within one call: both copies of a repeated position would target one device slot and
 * which one stays is undefined (hcm_host::MeshTaskDispatcher sends the latest object only).
 * Large uploads scan the host arrays for uniform chunks on worker threads (environment HCM_HOST_THREADS, default
 * min(15, cores - 1)); set it when several processes share one host, one ctx per GPU.
 * ---------------------------------------------------------------------------------------------------------- */
int hcm_upload_chunks(hcm_ctx *ctx, uint32_t n, const int16_t *pos /* n x 3 */, const uint16_t *blocks /* n x 32768 */,
                      const int8_t *heights /* n x 1024 */);
/* The same with what the caller already knows about each chunk, so that the host-side scan can be skipped (it reads every byte of a
 * uniform chunk: at r = 20 that is 1.9 GB of host DRAM traffic per full upload, the largest part of a world load's wall time):
 *   uniform_hint[i] >= 0              : every voxel of chunk i holds this Block value (the generator filled it with one block --
 *                                       client/src/ChunkGenerateTask.cpp:21-39 runs DefaultTerrain::Generate, which knows when a chunk
 *                                       is all air or all stone); the 64 KB are NOT read and not copied
 *   HCM_HINT_UNKNOWN     (-1)         : scan it (what hcm_upload_chunks does for every chunk)
 *   HCM_HINT_NOT_UNIFORM (-2)         : copy it without scanning
 * A wrong hint is the caller's bug: a chunk hinted uniform is derived from the hint, whatever its array holds.
 * uniform_hint == NULL is hcm_upload_chunks. */
#define HCM_HINT_UNKNOWN (-1)
#define HCM_HINT_NOT_UNIFORM (-2)
int hcm_upload_chunks_hinted(hcm_ctx *ctx, uint32_t n, const int16_t *pos /* n x 3 */, const uint16_t *blocks /* n x 32768 */,
                             const int8_t *heights /* n x 1024 */, const int32_t *uniform_hint /* n, or NULL */);
int hcm_remove_chunks(hcm_ctx *ctx, uint32_t n, const int16_t *pos /* n x 3 */);
/* Re-run the derivation kernel on chunks whose Block data is already device resident (n == 0: every resident chunk).
 * This is the device-side half of hcm_upload_chunks; a full remesh "with inputs resident in HBM" is
 * hcm_rederive_chunks(ctx, 0, NULL) followed by hcm_mesh_batch. */
int hcm_rederive_chunks(hcm_ctx *ctx, uint32_t n, const int16_t *pos /* n x 3 */);
int hcm_is_resident(const hcm_ctx *ctx, const int16_t pos[3]); /* 1 / 0 */
/* Halo replication between contexts ("if the pool is already device-resident, cudaMemcpyPeerAsync the boundary planes over
 * NVLink"): the Block data and sunlight heights of the listed chunks go from `src`'s resident world to `dst`'s, device to device
 * (cudaMemcpyPeerAsync: NVLink / NVSwitch between peer GPUs, a plain device copy when both contexts share a GPU), and are derived again
 * on `dst`.  With the world cut into x-slabs (one ctx per GPU, each holding its slab dilated by a chunk) this is how a chunk the owner
 * edited with hcm_set_blocks reaches the neighbours that hold it as halo -- no host round trip, no second PCIe crossing.
 * Every position must be resident in `src` (HCM_ERR_NOT_RESIDENT) and inside `dst`'s window (HCM_ERR_OUT_OF_WINDOW); positions must be
 * distinct.  Asynchronous like hcm_upload_chunks: the copies are ordered after everything enqueued on `src` before the call, later
 * work on `src` is ordered after them, later calls on `dst` see the chunks.  Both contexts must belong to the calling process; the
 * caller serialises calls on BOTH for the duration of this one. */
int hcm_copy_chunks_from(hcm_ctx *dst, hcm_ctx *src, uint32_t n, const int16_t *pos /* n x 3 */);

/* ------------------------------------------------------------------------------------------------------------
 * Block edits on the device-resident world: what the reference does in three task types,
 *   kSetBlock      ChunkTaskRunner<kSetBlock>::Run      client/src/ChunkSetBlockTask.cpp:38-129
 *   kFloodSunlight ChunkTaskRunner<kFloodSunlight>::Run client/src/ChunkFloodSunlightTask.cpp:33-66
 *   kSetSunlight   ChunkTaskRunner<kSetSunlight>::Run   client/src/ChunkSetSunlightTask.cpp:28-90
 * for ChunkSetBlockEntry deltas (common/include/common/Data.hpp:11-14), without a 64 KB re-upload per touched chunk:
 *   1. Chunk::SetBlock for every entry whose block differs from the resident one;
 *   2. the sunlight height of each edited column is re-derived in its chunk (32 if the chunk above is not lit down to
 *      its floor in that column, else one above the highest voxel without vertical light pass) and, while it changes,
 *      in the chunks below -- the kFloodSunlight -> kSetSunlight -> kFloodSunlight(below) chain, which stops where the
 *      chunk above a chunk is not resident (the reference keeps that task queued);
 *   3. the touched chunks go through the derivation kernel again;
 *   4. the mesh tasks the reference pushes are returned: per changed block the chunk itself and every neighbour i with
 *      ChunkShouldRemesh(block_pos - offset_i * 32) (ChunkSetBlockTask.cpp:100-107); per changed height the chunk and
 *      all 26 neighbours (ChunkSetSunlightTask.cpp:66-78 tests the column against the chunk itself for every i, which
 *      is always true).  Deduplicated; positions whose 27 chunks are not all resident are counted in n_pending instead.
 * The caller applies `sunlight` to its own Chunk objects (Chunk::SetSunlightHeight) and passes `remesh_pos` to
 * hcm_mesh_batch; `remesh_flags` tells the edit's own fan-out (high priority in the reference) from the positions only a
 * changed sunlight height reaches (low priority).  Pointers in the result belong to the ctx and stay valid until its next call.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct hcm_block_edit {
	int16_t chunk[3]; /* ChunkPos3 */
	uint16_t index;   /* InnerIndex3: x + (y*32 + z)*32 */
	uint16_t block;   /* Block::GetData() */
	uint16_t _pad;
} hcm_block_edit; /* 12 bytes */

typedef struct hcm_sunlight_change {
	int16_t chunk[3];
	uint16_t index; /* InnerIndex2: x + z*32 */
	int8_t old_height, new_height;
	uint16_t _pad;
} hcm_sunlight_change; /* 12 bytes */

typedef struct hcm_edit_result {
	uint32_t n_changed_blocks; /* entries whose block differed from the resident one */
	uint32_t n_sunlight;
	uint32_t n_remesh;
	uint32_t n_pending;
	const hcm_sunlight_change *sunlight; /* [n_sunlight], each (chunk, column) at most once */
	const int16_t *remesh_pos;           /* [n_remesh x 3] */
	const uint8_t *remesh_flags;         /* [n_remesh] HCM_REMESH_*: who pushes the task in the reference */
} hcm_edit_result;
#define HCM_REMESH_BLOCK_EDIT 1u /* Push<kMesh>(pos, high_priority = true), ChunkSetBlockTask.cpp:119 */
#define HCM_REMESH_SUNLIGHT 2u   /* Push<kMesh>(pos, false), ChunkSetSunlightTask.cpp:84 */

int hcm_set_blocks(hcm_ctx *ctx, uint32_t n, const hcm_block_edit *edits, hcm_edit_result *out);
int hcm_sync(hcm_ctx *ctx);

/* ------------------------------------------------------------------------------------------------------------
 * Meshing.  One task = one chunk position whose 26 neighbours are resident (the 27-array that
 * ChunkTaskData<kMesh>::Pop captures, ChunkMeshTask.cpp:14-39).  hcm_mesh_batch launches asynchronously;
 * hcm_mesh_fetch waits, copies the compacted output to pinned host memory owned by the ctx and returns views that
 * stay valid until the next hcm_mesh_batch / hcm_destroy.
 *
 * Output contract (client/include/client/BlockVertex.hpp:8-25): per task, the list of BlockMesh in the order
 * BlockMeshAlgo::Generate returns them (BlockMeshAlgo.hpp:494-521): meshes cut by the 16-bit index limit in
 * overflow order, then the opaque remainder, then the transparent remainder, empty ones skipped.  Zero meshes is a
 * valid result and must still be pushed (WorldRenderer.hpp:81-86).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct hcm_mesh_desc {
	uint32_t task;         /* index into the batch (the caller's task order, also after hcm_upload_and_mesh) */
	uint32_t transparent;  /* BlockMesh::transparent */
	uint64_t vertex_offset;/* first vertex in hcm_batch_view::vertices */
	uint64_t index_offset; /* first index in hcm_batch_view::indices */
	uint32_t vertex_count; /* BlockMesh::vertices.size() */
	uint32_t index_count;  /* BlockMesh::indices.size() */
	uint32_t aabb_min[3];  /* BlockMesh::aabb (AABB<uint32_t>, 1/16 voxel units; default min=UINT32_MAX,max=0 dep/AABB.hpp:11) */
	uint32_t aabb_max[3];
} hcm_mesh_desc; /* 56 bytes */

typedef struct hcm_task_result {
	uint32_t first_mesh; /* index into hcm_batch_view::meshes */
	uint32_t n_meshes;
} hcm_task_result;

typedef struct hcm_batch_view {
	uint32_t n_tasks;
	uint32_t n_meshes;
	uint64_t n_vertices;
	uint64_t n_indices;
	const hcm_task_result *tasks; /* [n_tasks] */
	const hcm_mesh_desc *meshes;  /* [n_meshes] */
	const uint64_t *vertices;     /* [n_vertices] packed BlockVertex {x10_y10_z10_axis2, tex10_trans3_face3_ao2_sl6_tl6} */
	const uint16_t *indices;      /* [n_indices] */
} hcm_batch_view;

int hcm_mesh_batch(hcm_ctx *ctx, uint32_t n, const int16_t *pos /* n x 3 */);
int hcm_mesh_fetch(hcm_ctx *ctx, hcm_batch_view *out);

/* Device-side completion only (no output copy): waits for the last batch and reports its totals. */
typedef struct hcm_batch_counts {
	uint32_t n_tasks, n_meshes;
	uint64_t n_vertices, n_indices;
	uint32_t n_worked; /* tasks that went through the full light+mesh kernel (not classified trivial) */
	uint32_t _pad;
} hcm_batch_counts;
int hcm_mesh_wait(hcm_ctx *ctx, hcm_batch_counts *out);

/* ------------------------------------------------------------------------------------------------------------
 * Device-side sink: meshes go straight into caller-owned DEVICE buffers, no vertex or index byte crosses the bus.
 * The reference renderer gives every BlockMesh a range of a MeshCluster's vertex and index buffer (MeshCluster::try_alloc,
 * client/include/client/mesh/MeshCluster.hpp:38-67: VMA virtual allocations, alignment sizeof(Vertex) / sizeof(Index)), then builds
 * two staging buffers per mesh and copies them in (mesh/MeshHandle.hpp:75-87, MeshCluster.hpp:76-118).  With the meshes already in
 * device memory the staging disappears:
 *   1. hcm_mesh_batch(ctx, n, pos)
 *   2. hcm_mesh_counts(ctx, &view)   waits and returns tasks + mesh descriptors (counts, AABBs, transparency) only;
 *                                    view.vertices and view.indices are NULL, view.n_vertices / n_indices are the totals
 *   3. the caller allocates a range per mesh, exactly as MeshCluster::try_alloc does, and fills one hcm_mesh_placement per mesh
 *   4. hcm_mesh_emit_into(...)       one kernel copies every placed mesh's vertices and indices from the ctx's arena to
 *                                    vertex_base + vertex_offset (in vertices) and index_base + index_offset (in indices).
 * vertex_base / index_base are device pointers valid on the ctx's device: cudaMalloc'd memory, or a Vulkan buffer imported with
 * cudaImportExternalMemory + cudaExternalMemoryGetMappedBuffer.  The call returns when the copies
 * are complete.  Ranges must not overlap; a range that does not fit its capacity is HCM_ERR_INVALID and nothing is written.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct hcm_mesh_placement {
	uint32_t mesh;          /* index into hcm_batch_view::meshes of the last batch */
	uint32_t _pad;
	uint64_t vertex_offset; /* first vertex of the mesh in the caller's vertex buffer (MeshInfo::m_vertex_offset) */
	uint64_t index_offset;  /* first index of the mesh in the caller's index buffer (MeshInfo::m_first_index) */
} hcm_mesh_placement; /* 24 bytes */

int hcm_mesh_counts(hcm_ctx *ctx, hcm_batch_view *out);
int hcm_mesh_emit_into(hcm_ctx *ctx, uint32_t n, const hcm_mesh_placement *placements, void *dev_vertex_base, uint64_t vertex_capacity /* vertices */,
                       void *dev_index_base, uint64_t index_capacity /* indices */);

/* World (re)load in one call: hcm_upload_chunks + hcm_mesh_batch + hcm_mesh_fetch with the three stages overlapped.
 * The reference streams in the same way: a mesh task is pushed as soon as the 27 chunks around it exist
 * (client/src/ChunkMeshTask.cpp:18-34, World::update pushes tasks per finished neighbour) and WorldRenderer takes each
 * mesh as it completes (ChunkMeshTask.cpp:66-74).  Here the upload is cut into blocks; after block b is on the device
 * the tasks whose neighbourhood is complete are meshed while block b+1 crosses the bus, and finished meshes return on
 * the second copy engine.  `out` is what hcm_mesh_fetch would return: out->tasks[i] belongs to task_pos[i]; mesh
 * bytes per task are identical to the unpipelined calls (mesh order inside the arrays may differ).  Every neighbour of
 * every task must be resident already or be part of this upload. */
int hcm_upload_and_mesh(hcm_ctx *ctx, uint32_t n_chunks, const int16_t *pos /* n_chunks x 3 */, const uint16_t *blocks,
                        const int8_t *heights, uint32_t n_tasks, const int16_t *task_pos /* n_tasks x 3 */, hcm_batch_view *out);

/* ------------------------------------------------------------------------------------------------------------
 * Measurement and debugging
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct hcm_stats {
	/* CUDA-event times of the most recent call of each kind, on the ctx stream, milliseconds */
	float upload_copy_ms;  /* host->device copy of the last hcm_upload_chunks */
	float derive_ms;       /* derive kernel of the last hcm_upload_chunks */
	float classify_ms;     /* classify kernel of the last hcm_mesh_batch */
	float mesh_ms;         /* mesh kernel of the last hcm_mesh_batch */
	float fetch_copy_ms;   /* device->host copies of the last hcm_mesh_fetch */
	uint32_t kernel_launches; /* kernels launched by this ctx since creation */
	uint64_t h2d_bytes;    /* cumulative */
	uint64_t d2h_bytes;    /* cumulative */
	uint64_t device_bytes; /* device memory held by the ctx */
	float sink_ms;         /* scatter kernel of the last hcm_mesh_emit_into */
	uint32_t upload_unscanned_chunks; /* chunks of the last upload the copy engine took before a scan thread did (sent as they are) */
	uint64_t p2p_bytes;    /* cumulative: bytes received device to device through hcm_copy_chunks_from */
} hcm_stats; /* 64 bytes */
int hcm_get_stats(hcm_ctx *ctx, hcm_stats *out);

void *hcm_stream(hcm_ctx *ctx); /* the ctx's cudaStream_t, for external event timing */

/* ------------------------------------------------------------------------------------------------------------
 * Kernel variants, qualified on the device that runs them
 * ------------------------------------------------------------------------------------------------------------
 * derive_kernel and mesh_kernel each exist in a baseline form and in variant forms that re-schedule the same arithmetic
 * (DESIGN.md section 3b).  A context starts on the baselines.  The first large calls (a re-derive / upload block of at least
 * `tune_min_chunks` chunks, a mesh batch of at least `tune_min_tasks` tasks) are A/B rounds: the baseline runs, then every
 * variant, then the baseline again (so that the call's result is the baseline's whatever happens); a hash kernel digests what
 * each run wrote -- per chunk the derived property bytes, light planes, heights and flags; per task the mesh descriptors,
 * vertex bytes and index bytes -- the variants' hashes are compared with the baseline's on the device, and every run is timed
 * with CUDA events.  After `tune_rounds` such calls the fastest variant whose every hash matched is adopted iff its time,
 * summed over the rounds, was at least 2 % below the baseline's; a variant with a single differing hash is never launched
 * again by this context.  Nothing here involves the host CPU in the computation: all forms are sm_100a kernels of this library.
 * Environment at hcm_create: HCM_TUNE=0 (baselines only), HCM_DERIVE_VARIANT / HCM_MESH_VARIANT = 0 .. n (forced). */
enum {
	HCM_TUNE_PENDING = 0,          /* not enough large calls yet: the baseline runs */
	HCM_TUNE_VARIANT_ADOPTED = 1,  /* identical bytes, faster: variant `selected` runs from now on */
	HCM_TUNE_BASELINE_KEPT = 2,    /* no variant with identical bytes was faster */
	HCM_TUNE_VARIANT_REJECTED = 3, /* baseline kept, and some variant's output differed from the baseline's on this device */
	HCM_TUNE_FORCED = 4,           /* selected by environment / hcm_set_option */
	HCM_TUNE_OFF = 5
};
#define HCM_MAX_VARIANTS 3
typedef struct hcm_kernel_tune {
	uint32_t selected;    /* 0 baseline, k = variant k: what the next launch uses */
	uint32_t state;       /* HCM_TUNE_* */
	uint32_t rounds;      /* A/B rounds completed */
	uint32_t n_variants;  /* variant forms this kernel has (<= HCM_MAX_VARIANTS) */
	float baseline_ms;    /* summed over the rounds (per round: the faster of the baseline's two runs); the rounds may differ in size */
	float variant_ms[HCM_MAX_VARIANTS];       /* summed over the rounds, same inputs */
	uint32_t mismatches[HCM_MAX_VARIANTS];    /* units (chunks / tasks) whose hashes differed, all rounds */
} hcm_kernel_tune; /* 44 bytes */
typedef struct hcm_tune_info {
	hcm_kernel_tune derive, mesh;
} hcm_tune_info; /* 88 bytes */
int hcm_get_tune_info(hcm_ctx *ctx, hcm_tune_info *out);

/* Run-time knobs of a context (what the HCM_* environment variables set at hcm_create).  Names: "upload_adaptive" (0 | 1),
 * "tune" (0 | 1), "derive_variant" / "mesh_variant" (0 .. n: forced; -1: back to qualification), "tune_min_chunks",
 * "tune_min_tasks", "tune_rounds"; "tune_fault" (test knob: the next A/B rounds see one corrupted hash, which must reject the
 * variant).  HCM_ERR_INVALID for an unknown name or a value out of range. */
int hcm_set_option(hcm_ctx *ctx, const char *name, int64_t value);

/* Light stage only, for stage-wise parity tests: the [-1,32]^3 light bytes the mesher reads
 * (index (x+1) + ((z+1) + (y+1)*34)*34; value = block::Light data = torch << 4 | sun,
 * block/include/block/Light.hpp:20-24) after the flood of ChunkMeshTask.cpp:48-111. */
int hcm_debug_light(hcm_ctx *ctx, const int16_t pos[3], uint8_t *out /* 39304 */);

#ifdef __cplusplus
}
#endif
#endif /* HCM_H */
