// TEST INFRASTRUCTURE (oracle build only) -- not part of the shipped product.
//
// Shadow of the reference header client/include/client/WorldRenderer.hpp, placed FIRST on the
// include path when oracle/_ref is built, so that the reference's unmodified
// client/src/ChunkMeshTask.cpp compiles without Vulkan/MyVK and so that
// `renderer->PushChunkMesh(pos, std::move(meshes))` (ChunkMeshTask.cpp:124-126) lands in a
// capture buffer instead of the Vulkan mesh pool (WorldRenderer.hpp:69-87).
//
// The class keeps the reference's name and namespace because World.hpp:37 declares
// `friend class WorldRenderer`, which is what lets Create() below set
// World::m_world_renderer_weak_ptr exactly like the real one does (WorldRenderer.hpp:41-46).
#pragma once

#include <client/BlockVertex.hpp>
#include <client/World.hpp>

#include <memory>
#include <vector>

namespace hc::client {

class WorldRenderer : public std::enable_shared_from_this<WorldRenderer> {
public:
	ChunkPos3 captured_pos{};
	std::vector<BlockMesh> captured;
	uint64_t push_count{0};

	static inline std::shared_ptr<WorldRenderer> Create(const std::shared_ptr<World> &world_ptr) {
		auto ret = std::make_shared<WorldRenderer>();
		world_ptr->m_world_renderer_weak_ptr = ret;
		return ret;
	}
	std::vector<ChunkPos3> *push_log{nullptr}; // when set: every pushed position, in order (the edit-chain harness of ref_shim.cpp)
	std::vector<std::vector<BlockMesh>> *mesh_log{nullptr};

	inline void PushChunkMesh(const ChunkPos3 &chunk_pos, std::vector<BlockMesh> &&meshes) {
		captured_pos = chunk_pos;
		if (push_log)
			push_log->push_back(chunk_pos);
		if (mesh_log)
			mesh_log->push_back(meshes);
		captured = std::move(meshes);
		++push_count;
	}
	// World::update() (client/src/World.cpp:7-12) tells the renderer to drop the meshes of unloaded chunks: nothing to drop here
	inline void EraseUnloadedMeshes() {}
};

} // namespace hc::client
