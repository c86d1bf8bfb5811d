// Drop-in replacement of the reference's client/src/ChunkMeshTask.cpp: the mesh lane of the chunk-task plugin interface
// (ChunkTaskData<kMesh>::Pop + ChunkTaskRunner<kMesh>::Run, declared in client/include/client/ChunkMeshTask.inl:17-64) on top of
// libhcm.so.  Compiled against the reference's own headers -- Chunk.hpp, ChunkTaskPool.hpp, World.hpp, WorldRenderer.hpp,
// BlockVertex.hpp -- with nothing of the reference modified; oracle/Makefile target `binding` builds it next to the unmodified
// rest of the client (test: tests/test_reference_binding.py).
//
//   reference                                                        here
//   Pop: readiness + capture of the 27 shared_ptr<Chunk>             the same rule (ChunkMeshTask.cpp:14-39)
//   Run on one of N worker threads: gather 62^3, flood, mesh,        Run = Submit to the dispatcher (a position asked for twice is
//   PushChunkMesh (ChunkMeshTask.cpp:41-127)                          one task); the first worker to find the ctx free drains
//                                                                    everything pending in ONE hcm_upload_and_mesh and pushes every
//                                                                    result to WorldRenderer::PushChunkMesh from its thread
#include <client/ChunkTaskPool.hpp>

#include <client/World.hpp>
#include <client/WorldRenderer.hpp>

#include <memory>
#include <mutex>
#include <unordered_map>

#include "hcm_binding.hpp"

namespace hcm_binding {
namespace {
std::mutex g_config_mutex;
std::shared_ptr<Dispatcher> g_dispatcher;
std::mutex g_flush_mutex; // held by the worker that drains the pending tasks

// priority of a popped task: Pop clears ChunkTaskData<kMesh>::m_high_priority and ChunkTaskRunnerData<kMesh> does not carry it
std::mutex g_priority_mutex;
std::unordered_map<hcm_host::ChunkPos3, bool, hcm_host::ChunkPos3Hash> g_priority;

std::shared_ptr<Dispatcher> dispatcher() {
	std::lock_guard<std::mutex> g(g_config_mutex);
	return g_dispatcher;
}
} // namespace

void Configure(const void *registry_blob, size_t blob_size, int device, hcm_host::ChunkPos3 window_origin, std::array<uint16_t, 3> window_extent) {
	auto d = std::make_shared<Dispatcher>(registry_blob, blob_size, device, window_origin, window_extent);
	std::lock_guard<std::mutex> g(g_config_mutex);
	g_dispatcher = std::move(d);
}
void Shutdown() {
	std::lock_guard<std::mutex> g(g_config_mutex);
	g_dispatcher.reset();
}
Dispatcher *Get() { return dispatcher().get(); }

} // namespace hcm_binding

namespace hc::client {

// Readiness as in the reference: queued, none of the 27 chunks still waiting for its generator, all 27 present and generated.
std::optional<ChunkTaskRunnerData<ChunkTaskType::kMesh>> ChunkTaskData<ChunkTaskType::kMesh>::Pop(const ChunkTaskPoolLocked &task_pool,
                                                                                                 const ChunkPos3 &chunk_pos) {
	if (!m_queued)
		return std::nullopt;
	std::array<std::shared_ptr<Chunk>, 27> chunks;
	for (uint32_t i = 0; i < 27; ++i) {
		ChunkPos3 offset;
		Chunk::NeighbourIndex2CmpXYZ(i, glm::value_ptr(offset));
		const ChunkPos3 pos = chunk_pos + offset;
		if (task_pool.AnyNotIdle<ChunkTaskType::kGenerate>(pos))
			return std::nullopt;
		chunks[i] = task_pool.GetWorld().GetChunkPool().FindChunk(pos);
		if (!chunks[i])
			return std::nullopt;
	}
	{
		std::lock_guard<std::mutex> g(hcm_binding::g_priority_mutex);
		hcm_binding::g_priority[hcm_host::ChunkPos3{chunk_pos.x, chunk_pos.y, chunk_pos.z}] = m_high_priority;
	}
	m_queued = false;
	m_high_priority = false;
	return ChunkTaskRunnerData<ChunkTaskType::kMesh>{std::move(chunks)};
}

void ChunkTaskRunner<ChunkTaskType::kMesh>::Run(ChunkTaskPool *p_task_pool, ChunkTaskRunnerData<ChunkTaskType::kMesh> &&data) {
	auto dispatcher = hcm_binding::dispatcher();
	if (!dispatcher)
		return; // not configured: like a missing renderer in the reference, the result is dropped (ChunkMeshTask.cpp:124-126)
	const ChunkPos3 &pos = data.GetChunkPos();
	bool high = false;
	{
		std::lock_guard<std::mutex> g(hcm_binding::g_priority_mutex);
		auto it = hcm_binding::g_priority.find(hcm_host::ChunkPos3{pos.x, pos.y, pos.z});
		if (it != hcm_binding::g_priority.end()) {
			high = it->second;
			hcm_binding::g_priority.erase(it);
		}
	}
	dispatcher->Submit(data.GetChunkPtrArray(), high);

	// Whoever finds the ctx free drains what is pending -- its own task and those the other workers submitted meanwhile.  A task
	// submitted while the drain's last flush is already running is taken by the re-check below or by the next Run().
	while (dispatcher->Pending() && hcm_binding::g_flush_mutex.try_lock()) {
		std::lock_guard<std::mutex> flush_lock(hcm_binding::g_flush_mutex, std::adopt_lock);
		auto renderer = p_task_pool->GetWorld().LockRenderer(); // the sink of ChunkMeshTask.cpp:124-126
		const size_t before = dispatcher->Pending();
		dispatcher->Flush([&](const hcm_host::ChunkPos3 &p, std::vector<hcm_host::BlockMesh> &&meshes) {
			if (!renderer)
				return;
			std::vector<BlockMesh> out(meshes.size()); // hc::client::BlockMesh, client/include/client/BlockVertex.hpp:20-25
			for (size_t i = 0; i < meshes.size(); ++i) {
				static_assert(sizeof(BlockVertex) == sizeof(hcm_host::BlockVertex), "8-byte packed vertex on both sides");
				const auto *v = reinterpret_cast<const BlockVertex *>(meshes[i].vertices.data());
				out[i].vertices.assign(v, v + meshes[i].vertices.size());
				out[i].indices = std::move(meshes[i].indices);
				auto &mn = meshes[i].aabb.min, &mx = meshes[i].aabb.max;
				if (mn[0] <= mx[0]) // a default AABB (min = UINT32_MAX, max = 0, dep/AABB.hpp:11) stays the default
					out[i].aabb = AABB<uint32_t>{{mn[0], mn[1], mn[2]}, {mx[0], mx[1], mx[2]}};
				out[i].transparent = meshes[i].transparent;
			}
			renderer->PushChunkMesh(ChunkPos3{p.x, p.y, p.z}, std::move(out));
		});
		if (dispatcher->Pending() >= before)
			break; // nothing could run (every pending task waits for an evicted neighbour): do not spin
	}
}

} // namespace hc::client
