// hcm_host.hpp -- C++ host side above the C ABI (include/hcm.h), mirroring the reference's mesh-task interface.
//
// The reference runs one `ChunkTaskRunner<ChunkTaskType::kMesh>::Run(pool, ChunkTaskRunnerData<kMesh>&&)` per chunk
// on N worker threads (client/src/ChunkMeshTask.cpp:41-127, client/include/client/ChunkTaskPool.hpp:187-194) and ends
// with `renderer->PushChunkMesh(chunk->GetPosition(), std::move(meshes))` (:124-126).  On a GPU the unit of dispatch is
// a BATCH of popped tasks: `MeshTaskDispatcher::Submit` takes the same 27-chunk array `ChunkTaskData<kMesh>::Pop`
// captures (ChunkMeshTask.cpp:14-39; index 26 = centre, order = NeighbourIndex2CmpXYZ, Position.hpp:146-152) and
// `Flush` hands every task's `std::vector<BlockMesh>` to a sink with PushChunkMesh's signature.
//
// `ChunkT` is any type with the reference Chunk's accessors (client/include/client/Chunk.hpp:45-66):
//     const uint16_t*-compatible  GetBlockData()      (Block is a 2-byte POD, Block.hpp:89)
//     int8_t                      GetSunlightHeight(uint32_t idx)
//     position with .x .y .z      GetPosition()
// so the reference's own `hc::client::Chunk` plugs in unchanged.
#pragma once

#include <array>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <exception>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/hcm.h"

namespace hcm_host {

// layout-compatible with hc::client::BlockVertex (client/include/client/BlockVertex.hpp:8-18)
struct BlockVertex {
	uint32_t x10_y10_z10_axis2, tex10_trans3_face3_ao2_sl6_tl6;
};
static_assert(sizeof(BlockVertex) == 8, "BlockVertex is 8 bytes (R32G32_UINT, rg/ChunkOpaquePass.hpp:70-71)");

struct U32AABB { // AABB<uint32_t>, dep/AABB.hpp:9-12
	uint32_t min[3]{UINT32_MAX, UINT32_MAX, UINT32_MAX}, max[3]{0, 0, 0};
};

// hc::client::BlockMesh (BlockVertex.hpp:20-25)
struct BlockMesh {
	std::vector<BlockVertex> vertices;
	std::vector<uint16_t> indices;
	U32AABB aabb{};
	bool transparent{};
};

struct ChunkPos3 {
	int16_t x, y, z;
	bool operator==(const ChunkPos3 &r) const { return x == r.x && y == r.y && z == r.z; }
};
struct ChunkPos3Hash {
	size_t operator()(const ChunkPos3 &p) const {
		return (size_t)(uint16_t)p.x * 73856093u ^ (size_t)(uint16_t)p.y * 19349663u ^ (size_t)(uint16_t)p.z * 83492791u;
	}
};

class Error : public std::runtime_error {
public:
	int status;
	Error(int s, const std::string &what) : std::runtime_error(what), status(s) {}
};

// The pending mesh requests of the world, one per chunk position -- what the per-chunk `ChunkTaskData<kMesh>` flags are in
// the reference: `Push(high_priority)` sets `m_queued` and ORs `m_high_priority` (ChunkMeshTask.inl:38-41), so a chunk asked
// for twice before it runs is meshed once, and an edit (`Push<kMesh>(pos, true)`, ChunkSetBlockTask.cpp:119) overtakes the
// streaming requests (`ChunkGenerateTask.cpp:38`, low).  `Pop(n)` is the producer's pick of a round
// (ChunkTaskPool.cpp:16-85): high priority first, each class in the order of first request.  No GPU involved.
class MeshRequestQueue {
public:
	// true if the position was not pending yet
	bool Push(const ChunkPos3 &pos, bool high_priority = false) {
		auto it = m_index.find(pos);
		if (it == m_index.end()) {
			m_index.emplace(pos, m_entries.size());
			m_entries.push_back(Entry{pos, high_priority, true});
			++m_live;
			return true;
		}
		m_entries[it->second].high |= high_priority;
		return false;
	}
	// ChunkTaskData<kMesh>::OnUnload (ChunkMeshTask.inl:48): the chunk left the world, its request with it
	bool Drop(const ChunkPos3 &pos) {
		auto it = m_index.find(pos);
		if (it == m_index.end())
			return false;
		m_entries[it->second].live = false;
		m_index.erase(it);
		--m_live;
		return true;
	}
	bool IsQueued(const ChunkPos3 &pos) const { return m_index.count(pos) != 0; }
	size_t Size() const { return m_live; }
	// at most max_tasks positions, removed from the queue
	std::vector<ChunkPos3> Pop(size_t max_tasks = SIZE_MAX, std::vector<bool> *high = nullptr) {
		std::vector<ChunkPos3> out;
		if (high)
			high->clear();
		for (int want_high = 1; want_high >= 0; --want_high)
			for (Entry &e : m_entries)
				if (e.live && e.high == (want_high != 0) && out.size() < max_tasks) {
					out.push_back(e.pos);
					if (high)
						high->push_back(e.high);
					e.live = false;
					m_index.erase(e.pos);
				}
		m_live -= out.size();
		std::vector<Entry> rest; // compact, keeping the order of first request
		for (const Entry &e : m_entries)
			if (e.live) {
				m_index[e.pos] = rest.size();
				rest.push_back(e);
			}
		m_entries.swap(rest);
		return out;
	}

private:
	struct Entry {
		ChunkPos3 pos;
		bool high, live;
	};
	std::vector<Entry> m_entries;
	std::unordered_map<ChunkPos3, size_t, ChunkPos3Hash> m_index;
	size_t m_live{0};
};

template <typename ChunkT> class MeshTaskDispatcher {
public:
	// registry_blob: the serialised hcm_registry (INTEGRATION.md shows the exporter); window: chunks the device may hold
	MeshTaskDispatcher(const void *registry_blob, size_t blob_size, int device, ChunkPos3 window_origin, std::array<uint16_t, 3> window_extent) {
		hcm_registry reg{};
		if (int rc = hcm_registry_from_blob(registry_blob, blob_size, &reg))
			throw Error(rc, "bad registry blob");
		hcm_config cfg{};
		cfg.device = device;
		cfg.origin[0] = window_origin.x, cfg.origin[1] = window_origin.y, cfg.origin[2] = window_origin.z;
		cfg.extent[0] = window_extent[0], cfg.extent[1] = window_extent[1], cfg.extent[2] = window_extent[2];
		if (int rc = hcm_create(&reg, &cfg, &m_ctx))
			throw Error(rc, hcm_last_error(nullptr));
	}
	~MeshTaskDispatcher() { hcm_destroy(m_ctx); }
	MeshTaskDispatcher(const MeshTaskDispatcher &) = delete;
	MeshTaskDispatcher &operator=(const MeshTaskDispatcher &) = delete;

	// A chunk whose blocks or sunlight heights changed since it was last seen (the SetBlock / SetSunlight tasks that
	// `Push<kMesh>`, ChunkSetBlockTask.cpp:119, ChunkSetSunlightTask.cpp:84) must be re-uploaded before its remesh.
	void MarkDirty(const ChunkPos3 &pos) {
		std::lock_guard<std::mutex> g(m_mutex);
		m_resident.erase(pos);
	}

	// Same input as ChunkTaskRunnerData<kMesh> (ChunkMeshTask.inl:17-29).  Like Run(), the chunks are read-only snapshots
	// held by shared_ptr until Flush() returns.
	// A position submitted again before it was flushed stays ONE task (the later snapshot wins, the priority is sticky):
	// MeshRequestQueue above.  high_priority = the `true` of Push<kMesh>(pos, true) that edits use.
	// The chunk left the world (ChunkPool::Update evicts it, client/src/ChunkPool.cpp:9-33, and the task pool calls
	// ChunkTaskData<kMesh>::OnUnload, ChunkMeshTask.inl:48): its device slot is freed, its pending request and its not yet
	// uploaded copy are dropped.  Requests of its neighbours stay pending: like Pop() in the reference (ChunkMeshTask.cpp:29-31
	// returns nullopt for that one task) Flush leaves them out of its round until the chunk has come back.
	// Lock order as in Flush -- ctx first, then the bookkeeping -- and both held to the end: a Submit + Flush of a new object for
	// `pos` cannot slip between the bookkeeping and hcm_remove_chunks (which would then delete the fresh device copy).
	void Unload(const ChunkPos3 &pos) {
		std::lock_guard<std::mutex> ctx_lock(m_ctx_mutex);
		std::lock_guard<std::mutex> g(m_mutex);
		m_resident.erase(pos);
		m_queue.Drop(pos);
		m_snapshot.erase(pos);
		m_check_readiness = true; // a pending task may now lack one of its 27 chunks
		for (size_t i = 0; i < m_upload.size();) {
			const auto p = m_upload[i]->GetPosition();
			if (ChunkPos3{(int16_t)p.x, (int16_t)p.y, (int16_t)p.z} == pos)
				m_upload.erase(m_upload.begin() + (std::ptrdiff_t)i);
			else
				++i;
		}
		const int16_t p[3] = {pos.x, pos.y, pos.z};
		check(hcm_remove_chunks(m_ctx, 1, p));
	}

	// Thread-safe like Run(): the reference calls it from N worker threads (ChunkTaskPool.cpp:119-129); Submit / MarkDirty only
	// touch the bookkeeping, the calls into the ctx (Flush, SetBlocks) are serialised among themselves.
	void Submit(const std::array<std::shared_ptr<ChunkT>, 27> &chunks, bool high_priority = false) {
		std::lock_guard<std::mutex> g(m_mutex);
		for (const auto &c : chunks) {
			const auto p = c->GetPosition();
			ChunkPos3 pos{(int16_t)p.x, (int16_t)p.y, (int16_t)p.z};
			auto it = m_resident.find(pos);
			if (it == m_resident.end() || it->second != c.get()) {
				m_resident[pos] = c.get();
				m_upload.push_back(c);
			}
		}
		const auto p = chunks[26]->GetPosition();
		const ChunkPos3 centre{(int16_t)p.x, (int16_t)p.y, (int16_t)p.z};
		m_queue.Push(centre, high_priority);
		m_snapshot[centre] = chunks;
	}

	size_t Pending() const {
		std::lock_guard<std::mutex> g(m_mutex);
		return m_queue.Size();
	}

	// Halo replication: chunks that are resident on `src`'s GPU become resident here, device to device
	// (hcm_copy_chunks_from: cudaMemcpyPeerAsync, NVLink between peer GPUs) -- the way a chunk the owner of a slab edited on its device
	// (SetBlocks) reaches the neighbour that holds it as halo, without the 64 KB going through the host again.  The bookkeeping
	// follows: this dispatcher now knows the chunk OBJECT `src` knows (the host edited that object in place, so it describes the
	// device copy), and a copy of the position that was still waiting to go up from the host is dropped.
	void CopyChunksFrom(MeshTaskDispatcher &src, const std::vector<ChunkPos3> &positions) {
		if (positions.empty() || &src == this)
			return;
		// both contexts are busy for the duration of the call; lock in address order (two dispatchers may copy from each other)
		std::mutex &first = this < &src ? m_ctx_mutex : src.m_ctx_mutex, &second = this < &src ? src.m_ctx_mutex : m_ctx_mutex;
		std::lock_guard<std::mutex> l1(first), l2(second);
		std::vector<int16_t> p(positions.size() * 3);
		for (size_t i = 0; i < positions.size(); ++i)
			p[3 * i] = positions[i].x, p[3 * i + 1] = positions[i].y, p[3 * i + 2] = positions[i].z;
		check(hcm_copy_chunks_from(m_ctx, src.m_ctx, (uint32_t)positions.size(), p.data()));
		std::unordered_map<ChunkPos3, const ChunkT *, ChunkPos3Hash> known;
		{
			std::lock_guard<std::mutex> g(src.m_mutex);
			for (const ChunkPos3 &q : positions) {
				auto it = src.m_resident.find(q);
				if (it != src.m_resident.end())
					known.emplace(q, it->second);
			}
		}
		std::lock_guard<std::mutex> g(m_mutex);
		for (const ChunkPos3 &q : positions) {
			auto it = known.find(q);
			if (it != known.end())
				m_resident[q] = it->second;
			else
				m_resident.erase(q); // nobody knows which object this is: the next task that reads it uploads the host's copy
			for (size_t i = 0; i < m_upload.size();) {
				const auto up = m_upload[i]->GetPosition();
				if (ChunkPos3{(int16_t)up.x, (int16_t)up.y, (int16_t)up.z} == q && (it == known.end() || m_upload[i].get() != it->second))
					m_upload.erase(m_upload.begin() + (std::ptrdiff_t)i);
				else
					++i;
			}
		}
	}

	// ChunkSetBlockEntry deltas (common/include/common/Data.hpp:11-14) for chunks that are already on the device -- the
	// work of ChunkTaskRunner<kSetBlock>::Run and of the kFloodSunlight / kSetSunlight chain it starts
	// (client/src/ChunkSetBlockTask.cpp:78-129, ChunkFloodSunlightTask.cpp:41-57, ChunkSetSunlightTask.cpp:42-90) without
	// re-uploading 64 KB per touched chunk.  The caller has already applied (or now applies) the same blocks to its own
	// Chunk objects with Chunk::SetBlock; `on_sunlight(const SunlightChange&)` is where it calls Chunk::SetSunlightHeight,
	// `on_remesh(const ChunkPos3&, bool high_priority)` where it would Push<kMesh> -- typically Submit(the 27 chunks around that
	// position, high_priority).  high_priority is the reference's: true for the fan-out of a changed block
	// (ChunkSetBlockTask.cpp:119), false for positions that only a changed sunlight height reaches (ChunkSetSunlightTask.cpp:84).
	// The callbacks run inside the ctx's critical section: Submit / MarkDirty are fine there, Flush / SetBlocks are not.
	struct SetBlock {
		ChunkPos3 chunk;
		uint16_t index; // InnerIndex3: x + (y*32 + z)*32
		uint16_t block; // Block::GetData()
	};
	struct SunlightChange {
		ChunkPos3 chunk;
		uint16_t index; // InnerIndex2: x + z*32
		int8_t old_height, new_height;
	};
	template <typename OnSunlight, typename OnRemesh> uint32_t SetBlocks(const std::vector<SetBlock> &edits, OnSunlight &&on_sunlight, OnRemesh &&on_remesh) {
		std::vector<hcm_block_edit> e(edits.size());
		for (size_t i = 0; i < edits.size(); ++i) {
			e[i].chunk[0] = edits[i].chunk.x, e[i].chunk[1] = edits[i].chunk.y, e[i].chunk[2] = edits[i].chunk.z;
			e[i].index = edits[i].index, e[i].block = edits[i].block, e[i]._pad = 0;
		}
		std::lock_guard<std::mutex> ctx_lock(m_ctx_mutex); // the result's pointers belong to the ctx until its next call
		hcm_edit_result r{};
		check(hcm_set_blocks(m_ctx, (uint32_t)e.size(), e.data(), &r));
		for (uint32_t i = 0; i < r.n_sunlight; ++i) {
			const hcm_sunlight_change &s = r.sunlight[i];
			on_sunlight(SunlightChange{ChunkPos3{s.chunk[0], s.chunk[1], s.chunk[2]}, s.index, s.old_height, s.new_height});
		}
		for (uint32_t i = 0; i < r.n_remesh; ++i)
			on_remesh(ChunkPos3{r.remesh_pos[3 * i], r.remesh_pos[3 * i + 1], r.remesh_pos[3 * i + 2]}, (r.remesh_flags[i] & HCM_REMESH_BLOCK_EDIT) != 0);
		return r.n_changed_blocks;
	}

	// Runs the batch and calls sink(const ChunkPos3&, std::vector<BlockMesh>&&) once per submitted task, in submission
	// order -- the place where Run() calls WorldRenderer::PushChunkMesh.  Zero meshes is a valid result and is pushed too
	// (it clears the chunk's old handles, WorldRenderer.hpp:81-86).
	// max_tasks: the budget of one round (the reference producer hands out at most 256 tasks per round,
	// ChunkTaskPool.cpp:16-85); what does not fit stays pending, high priority goes first.
	template <typename Sink> void Flush(Sink &&sink, size_t max_tasks = SIZE_MAX) {
		std::lock_guard<std::mutex> ctx_lock(m_ctx_mutex); // one call into the ctx at a time; Submit stays open meanwhile
		std::vector<ChunkPos3> tasks;
		std::vector<bool> high;
		std::vector<std::shared_ptr<ChunkT>> upload;
		std::vector<std::array<std::shared_ptr<ChunkT>, 27>> snapshots; // read-only inputs, held until the sink has run
		bool check_readiness;
		{
			std::lock_guard<std::mutex> g(m_mutex);
			check_readiness = m_check_readiness;
			tasks = m_queue.Pop(max_tasks, &high);
			if (tasks.empty())
				return;
			upload.swap(m_upload);
			for (const ChunkPos3 &t : tasks) {
				auto it = m_snapshot.find(t);
				snapshots.push_back(std::move(it->second));
				m_snapshot.erase(it);
			}
		}
		if (upload.size() > 1) { // a chunk that was replaced while its first copy was still waiting: only the latest object goes up
			std::unordered_set<ChunkPos3, ChunkPos3Hash> seen;
			std::vector<std::shared_ptr<ChunkT>> latest;
			for (size_t i = upload.size(); i-- > 0;) {
				const auto p = upload[i]->GetPosition();
				if (seen.insert(ChunkPos3{(int16_t)p.x, (int16_t)p.y, (int16_t)p.z}).second)
					latest.push_back(std::move(upload[i]));
			}
			upload.assign(latest.rbegin(), latest.rend());
		}
		const size_t n = upload.size();
		if (n) {
			m_pos.resize(n * 3), m_blocks.resize(n * HCM_CHUNK_VOXELS), m_heights.resize(n * HCM_CHUNK_COLUMNS);
			for (size_t i = 0; i < n; ++i) {
				const auto p = upload[i]->GetPosition();
				m_pos[3 * i] = (int16_t)p.x, m_pos[3 * i + 1] = (int16_t)p.y, m_pos[3 * i + 2] = (int16_t)p.z;
				std::memcpy(m_blocks.data() + i * HCM_CHUNK_VOXELS, upload[i]->GetBlockData(), HCM_CHUNK_VOXELS * 2);
				for (uint32_t k = 0; k < HCM_CHUNK_COLUMNS; ++k)
					m_heights[i * HCM_CHUNK_COLUMNS + k] = (int8_t)upload[i]->GetSunlightHeight(k);
			}
		}
		// Readiness per task, as ChunkTaskData<kMesh>::Pop tests it (ChunkMeshTask.cpp:18-34): a task one of whose 27 chunks is
		// neither on the device nor part of this upload (its neighbour was Unloaded after the request) is left out of the round
		// and stays pending; every other task runs.  One evicted chunk must not stall the world.  Submit names all 27 chunks of a
		// task and queues the ones the device lacks, so only an Unload can open such a gap: the test (27 look-ups per task) runs only
		// after one, and until no pending task waits for a chunk any more.
		if (check_readiness) {
			std::unordered_set<ChunkPos3, ChunkPos3Hash> coming;
			for (const auto &c : upload) {
				const auto p = c->GetPosition();
				coming.insert(ChunkPos3{(int16_t)p.x, (int16_t)p.y, (int16_t)p.z});
			}
			std::vector<size_t> deferred;
			for (size_t i = 0; i < tasks.size(); ++i) {
				bool ready = true;
				for (int dx = -1; dx <= 1 && ready; ++dx)
					for (int dy = -1; dy <= 1 && ready; ++dy)
						for (int dz = -1; dz <= 1 && ready; ++dz) {
							const ChunkPos3 q{(int16_t)(tasks[i].x + dx), (int16_t)(tasks[i].y + dy), (int16_t)(tasks[i].z + dz)};
							const int16_t qp[3] = {q.x, q.y, q.z};
							ready = coming.count(q) || hcm_is_resident(m_ctx, qp);
						}
				if (!ready)
					deferred.push_back(i);
			}
			if (deferred.empty()) {
				std::lock_guard<std::mutex> g(m_mutex);
				if (m_queue.Size() == 0)
					m_check_readiness = false; // every pending task was looked at and is ready
			} else {
				std::lock_guard<std::mutex> g(m_mutex);
				for (size_t k = deferred.size(); k-- > 0;) {
					const size_t i = deferred[k];
					m_queue.Push(tasks[i], high[i]);
					m_snapshot.emplace(tasks[i], std::move(snapshots[i])); // a newer snapshot, if one came in, wins
					tasks.erase(tasks.begin() + (std::ptrdiff_t)i), high.erase(high.begin() + (std::ptrdiff_t)i);
					snapshots.erase(snapshots.begin() + (std::ptrdiff_t)i);
				}
				if (tasks.empty()) { // nothing can run: the chunks waiting to go up wait with their tasks
					m_upload.insert(m_upload.begin(), upload.begin(), upload.end());
					return;
				}
			}
		}
		std::vector<int16_t> tp(tasks.size() * 3);
		for (size_t i = 0; i < tasks.size(); ++i)
			tp[3 * i] = tasks[i].x, tp[3 * i + 1] = tasks[i].y, tp[3 * i + 2] = tasks[i].z;
		// one call: new chunks go up, every task is meshed as soon as the last of its 27 chunks is on the device, finished
		// meshes come back while later chunks are still uploading (with nothing to upload it is batch + fetch)
		hcm_batch_view v{};
		if (int rc = hcm_upload_and_mesh(m_ctx, (uint32_t)n, m_pos.data(), m_blocks.data(), m_heights.data(), (uint32_t)tasks.size(), tp.data(), &v)) {
			// nothing was pushed: the round goes back to pending, behind the requests that arrived meanwhile
			std::lock_guard<std::mutex> g(m_mutex);
			m_upload.insert(m_upload.begin(), upload.begin(), upload.end());
			requeue(tasks, high, snapshots, 0);
			throw Error(rc, hcm_last_error(m_ctx));
		}
		uint32_t t = 0;
		try {
			push_results(v, tasks, t, sink);
		} catch (...) { // the sink threw at task t: that task and the ones behind it were not delivered -- they stay pending
			std::lock_guard<std::mutex> g(m_mutex);
			requeue(tasks, high, snapshots, t);
			throw;
		}
	}

	hcm_ctx *Context() { return m_ctx; }

private:
	void requeue(std::vector<ChunkPos3> &tasks, const std::vector<bool> &high, std::vector<std::array<std::shared_ptr<ChunkT>, 27>> &snapshots, size_t from) {
		for (size_t i = from; i < tasks.size(); ++i) {
			m_queue.Push(tasks[i], high[i]);
			m_snapshot.emplace(tasks[i], std::move(snapshots[i])); // a newer snapshot, if one came in, wins
		}
	}
	template <typename Sink> void push_results(const hcm_batch_view &v, const std::vector<ChunkPos3> &tasks, uint32_t &t, Sink &&sink) {
		for (t = 0; t < v.n_tasks; ++t) {
			std::vector<BlockMesh> meshes(v.tasks[t].n_meshes);
			for (uint32_t k = 0; k < v.tasks[t].n_meshes; ++k) {
				const hcm_mesh_desc &d = v.meshes[v.tasks[t].first_mesh + k];
				BlockMesh &m = meshes[k];
				m.vertices.resize(d.vertex_count);
				std::memcpy(m.vertices.data(), v.vertices + d.vertex_offset, (size_t)d.vertex_count * 8);
				m.indices.assign(v.indices + d.index_offset, v.indices + d.index_offset + d.index_count);
				std::memcpy(m.aabb.min, d.aabb_min, 12), std::memcpy(m.aabb.max, d.aabb_max, 12);
				m.transparent = d.transparent != 0;
			}
			sink(tasks[t], std::move(meshes));
		}
	}
	void check(int rc) {
		if (rc)
			throw Error(rc, hcm_last_error(m_ctx));
	}
	hcm_ctx *m_ctx{};
	mutable std::mutex m_mutex; // m_resident, m_upload, m_queue, m_snapshot
	std::mutex m_ctx_mutex;     // calls into m_ctx and the staging vectors below
	std::unordered_map<ChunkPos3, const ChunkT *, ChunkPos3Hash> m_resident;
	std::vector<std::shared_ptr<ChunkT>> m_upload;
	MeshRequestQueue m_queue;
	bool m_check_readiness{false}; // an Unload happened: Flush tests its tasks' 27 chunks until nothing pending waits for one
	std::unordered_map<ChunkPos3, std::array<std::shared_ptr<ChunkT>, 27>, ChunkPos3Hash> m_snapshot; // held until the task is flushed
	std::vector<int16_t> m_pos;
	std::vector<uint16_t> m_blocks;
	std::vector<int8_t> m_heights;
};

// One MeshTaskDispatcher per GPU over x-slabs of the world: chunks are independent given their halo, so
// meshing needs no exchange between the GPUs -- a task runs on the GPU whose slab [x_lo, x_hi) holds its position, and that GPU's
// window spans the planes x_lo - 1 .. x_hi, i.e. the boundary planes live on both neighbours.  The host's Chunk objects
// are the authority: a chunk that changes on the host is marked dirty on every GPU that can hold it and goes up again with the
// next task that reads it.  Edits (SetBlocks) run on the owner of the edited chunk, whose column of the world is complete
// (the sunlight chain walks along y, the cut is along x); GPUs holding a touched chunk as halo do not run the chain, they
// take the chunk as data -- from the owner's GPU, device to device (the one exchange step of the path: hcm_copy_chunks_from,
// cudaMemcpyPeerAsync over NVLink), or, with peer_halo = false, from the host.  hypercraft_b200/shard.py is the same bookkeeping
// for the Python benchmark.
template <typename ChunkT> class ShardedMeshDispatcher {
public:
	using Shard = MeshTaskDispatcher<ChunkT>;
	struct Slab {
		int device;        // CUDA ordinal
		int16_t x_lo, x_hi; // tasks with x_lo <= x < x_hi
	};
	// world window: origin / extent in y and z are shared, x comes from the slabs (which must tile an x-range, in order)
	// peer_halo: a chunk the owner edited on its device reaches the GPUs that hold it as halo device to device
	// (MeshTaskDispatcher::CopyChunksFrom); false: those GPUs are told the chunk is dirty and take the host's copy with their next task.
	ShardedMeshDispatcher(const void *registry_blob, size_t blob_size, const std::vector<Slab> &slabs, int16_t origin_y, int16_t origin_z,
	                      uint16_t extent_y, uint16_t extent_z, bool peer_halo = true)
	    : m_slabs(slabs), m_peer_halo(peer_halo) {
		if (slabs.empty())
			throw Error(HCM_ERR_INVALID, "no slabs");
		for (size_t r = 0; r < slabs.size(); ++r) {
			if (slabs[r].x_hi <= slabs[r].x_lo || (r && slabs[r].x_lo != slabs[r - 1].x_hi))
				throw Error(HCM_ERR_INVALID, "slabs must be non-empty and tile the x-range in order");
			m_shards.push_back(std::make_unique<Shard>(registry_blob, blob_size, slabs[r].device, ChunkPos3{(int16_t)(slabs[r].x_lo - 1), origin_y, origin_z},
			                                           std::array<uint16_t, 3>{(uint16_t)(slabs[r].x_hi - slabs[r].x_lo + 2), extent_y, extent_z}));
		}
	}

	size_t Ranks() const { return m_shards.size(); }
	Shard &Rank(size_t r) { return *m_shards[r]; }
	// rank whose slab holds x; -1 outside all slabs
	int Owner(int x) const {
		for (size_t r = 0; r < m_slabs.size(); ++r)
			if (x >= m_slabs[r].x_lo && x < m_slabs[r].x_hi)
				return (int)r;
		return -1;
	}
	// ranks whose window (slab dilated by one plane) reaches the x-plane, the owner included
	std::vector<int> Holders(int x) const {
		std::vector<int> out;
		for (size_t r = 0; r < m_slabs.size(); ++r)
			if (x >= m_slabs[r].x_lo - 1 && x <= m_slabs[r].x_hi)
				out.push_back((int)r);
		return out;
	}

	void Submit(const std::array<std::shared_ptr<ChunkT>, 27> &chunks, bool high_priority = false) {
		m_shards[owner_or_throw((int)chunks[26]->GetPosition().x)]->Submit(chunks, high_priority);
	}
	void MarkDirty(const ChunkPos3 &pos) {
		for (int r : Holders(pos.x))
			m_shards[r]->MarkDirty(pos);
	}
	void Unload(const ChunkPos3 &pos) {
		for (int r : Holders(pos.x))
			m_shards[r]->Unload(pos);
	}
	size_t Pending() const {
		size_t n = 0;
		for (const auto &s : m_shards)
			n += s->Pending();
		return n;
	}

	// ChunkShouldRemesh (client/include/client/ChunkMeshTask.inl:9-15): a block at `rel` (relative to a chunk's origin) can change
	// that chunk's mesh iff its L1 distance to the box [-1,33]^3 is below 15 (BlockLightAlgo.hpp:32-42 with level 15)
	static bool ShouldRemesh(int rx, int ry, int rz) {
		int dist = 0;
		for (int c : {rx, ry, rz})
			dist += c < -1 ? -1 - c : (c >= HCM_CHUNK_SIZE + 1 ? c - (HCM_CHUNK_SIZE + 1) : 0);
		return dist < 15;
	}

	// Edits go to the owner of their chunk; its heights and its remesh fan-out are the undivided world's for every position it
	// can see whole.  A fan-out position in ANOTHER slab is not reported by the owner's ctx (one of its 27 chunks lies outside
	// that window), so those are derived here from the reference's rules: per edit the neighbours passing ChunkShouldRemesh
	// (ChunkSetBlockTask.cpp:100-107; every edit is taken as a change -- a superset by the no-op edits), per changed height
	// the 26 neighbours (ChunkSetSunlightTask.cpp:66-78).  `on_sunlight` / `on_remesh` as in MeshTaskDispatcher::SetBlocks: the
	// caller updates its Chunk objects in on_sunlight and answers on_remesh -- for positions whose 27 chunks exist, the test
	// Pop() makes in the reference -- with Submit(the 27 chunks, true), which lands on the rank that owns the position.  Every
	// touched chunk is marked dirty on the ranks that hold it as halo: the next task there that reads it uploads the host's
	// (by then current) copy.  on_remesh runs after every owner has applied its edits.
	using SetBlock = typename Shard::SetBlock;
	using SunlightChange = typename Shard::SunlightChange;
	template <typename OnSunlight, typename OnRemesh> uint32_t SetBlocks(const std::vector<SetBlock> &edits, OnSunlight &&on_sunlight, OnRemesh &&on_remesh) {
		std::vector<std::vector<SetBlock>> by_rank(m_shards.size());
		for (const SetBlock &e : edits)
			by_rank[owner_or_throw(e.chunk.x)].push_back(e);
		uint32_t changed = 0;
		std::vector<ChunkPos3> remesh;
		std::unordered_map<ChunkPos3, bool, ChunkPos3Hash> seen; // position -> high priority (sticky, as in MeshRequestQueue)
		auto add = [&](const ChunkPos3 &p, bool high) {
			auto ins = seen.emplace(p, high);
			if (ins.second)
				remesh.push_back(p);
			else
				ins.first->second = ins.first->second || high;
		};
		for (size_t r = 0; r < m_shards.size(); ++r) {
			if (by_rank[r].empty())
				continue;
			auto add_foreign = [&](const ChunkPos3 &c, int dx, int dy, int dz, bool high) { // a fan-out position the owner's ctx cannot report
				const ChunkPos3 q{(int16_t)(c.x + dx), (int16_t)(c.y + dy), (int16_t)(c.z + dz)};
				const int o = Owner(q.x);
				if (o >= 0 && o != (int)r)
					add(q, high);
			};
			std::vector<ChunkPos3> touched;
			for (const SetBlock &e : by_rank[r]) {
				touched.push_back(e.chunk);
				const int bx = e.index & 31, by = e.index >> 10, bz = (e.index >> 5) & 31; // InnerIndex3: x + (y*32 + z)*32
				for (int dx = -1; dx <= 1; dx += 2) // only a step in x leaves the slab
					for (int dy = -1; dy <= 1; ++dy)
						for (int dz = -1; dz <= 1; ++dz)
							if (ShouldRemesh(bx - dx * HCM_CHUNK_SIZE, by - dy * HCM_CHUNK_SIZE, bz - dz * HCM_CHUNK_SIZE))
								add_foreign(e.chunk, dx, dy, dz, true);
			}
			changed += m_shards[r]->SetBlocks(
			    by_rank[r],
			    [&](const SunlightChange &s) {
				    touched.push_back(s.chunk);
				    for (int dx = -1; dx <= 1; dx += 2)
					    for (int dy = -1; dy <= 1; ++dy)
						    for (int dz = -1; dz <= 1; ++dz)
							    add_foreign(s.chunk, dx, dy, dz, false);
				    on_sunlight(s);
			    },
			    [&](const ChunkPos3 &p, bool high) { add(p, high); });
			std::vector<std::vector<ChunkPos3>> to_holder(m_shards.size());
			std::unordered_set<ChunkPos3, ChunkPos3Hash> once;
			for (const ChunkPos3 &c : touched)
				if (once.insert(c).second)
					for (int h : Holders(c.x))
						if (h != (int)r)
							to_holder[(size_t)h].push_back(c);
			for (size_t h = 0; h < m_shards.size(); ++h) {
				if (to_holder[h].empty())
					continue;
				if (m_peer_halo)
					m_shards[h]->CopyChunksFrom(*m_shards[r], to_holder[h]); // ordered after the owner's edit kernels (same stream)
				else
					for (const ChunkPos3 &c : to_holder[h])
						m_shards[h]->MarkDirty(c);
			}
		}
		for (const ChunkPos3 &p : remesh)
			on_remesh(p, seen[p]);
		return changed;
	}

	// Every rank's round, one after the other on the calling thread (a HyperCraft build would give each rank its own thread:
	// the shards share nothing).  max_tasks is per rank.
	template <typename Sink> void Flush(Sink &&sink, size_t max_tasks = SIZE_MAX) {
		for (auto &s : m_shards)
			s->Flush(sink, max_tasks);
	}

	// The same with one host thread per GPU, so that the GPUs work at the same time.  `sink` is called from those threads
	// concurrently -- as WorldRenderer::PushChunkMesh is in the reference, where every worker pushes its own result (it is
	// thread-safe: cuckoo map + MPMC queue, WorldRenderer.hpp:69-87).  The first error of any rank is rethrown after all ranks
	// have finished; the other ranks' rounds are not affected (a failed round stays pending on its rank).
	template <typename Sink> void FlushParallel(Sink &&sink, size_t max_tasks = SIZE_MAX) {
		std::vector<std::thread> threads;
		std::vector<std::exception_ptr> errors(m_shards.size());
		for (size_t r = 0; r < m_shards.size(); ++r)
			threads.emplace_back([&, r] {
				try {
					m_shards[r]->Flush(sink, max_tasks);
				} catch (...) {
					errors[r] = std::current_exception();
				}
			});
		for (auto &t : threads)
			t.join();
		for (auto &e : errors)
			if (e)
				std::rethrow_exception(e);
	}

private:
	size_t owner_or_throw(int x) const {
		const int r = Owner(x);
		if (r < 0)
			throw Error(HCM_ERR_OUT_OF_WINDOW, "position outside every slab");
		return (size_t)r;
	}
	std::vector<Slab> m_slabs;
	bool m_peer_halo;
	std::vector<std::unique_ptr<Shard>> m_shards;
};

} // namespace hcm_host
