// hcm_binding.hpp -- the process-wide mesh dispatcher of a HyperCraft client built with the B200 mesher.
//
// A HyperCraft build replaces client/src/ChunkMeshTask.cpp by ChunkMeshTask.cpp of this directory and calls
// hcm_binding::Configure once at start-up, next to WorldWorker::Launch (client/src/Application.cpp:248), with the registry blob
//  and the chunk window the device may hold (load sphere + unload margin around the spawn).
// Everything else -- Chunk, ChunkTaskPool, World, WorldRenderer, the shaders -- stays as it is.
#pragma once

#include <client/Chunk.hpp>

#include "../hcm_host.hpp"

namespace hcm_binding {

using Dispatcher = hcm_host::MeshTaskDispatcher<hc::client::Chunk>;

// Creates (or replaces) the dispatcher.  Throws hcm_host::Error when the registry cannot be mirrored or no sm_100 device is usable.
void Configure(const void *registry_blob, size_t blob_size, int device, hcm_host::ChunkPos3 window_origin, std::array<uint16_t, 3> window_extent);
void Shutdown();
Dispatcher *Get(); // nullptr before Configure

// The two places of the reference that change a chunk after it was generated -- ChunkTaskRunner<kSetBlock>::Run
// (client/src/ChunkSetBlockTask.cpp:92) and ChunkTaskRunner<kSetSunlight>::Run (client/src/ChunkSetSunlightTask.cpp:58) -- call
// this before their Push<kMesh>, unless they go through Dispatcher::SetBlocks, which edits on the device.
inline void MarkDirty(const hc::ChunkPos3 &pos) {
	if (Dispatcher *d = Get())
		d->MarkDirty(hcm_host::ChunkPos3{pos.x, pos.y, pos.z});
}

// ChunkPool::Update erases the chunks beyond the unload radius (client/src/ChunkPool.cpp:17-23): the same loop calls this, so that the
// device slot is freed and the chunk's pending request and not yet uploaded copy are dropped (ChunkTaskData<kMesh>::OnUnload).
inline void Unload(const hc::ChunkPos3 &pos) {
	if (Dispatcher *d = Get())
		d->Unload(hcm_host::ChunkPos3{pos.x, pos.y, pos.z});
}

} // namespace hcm_binding
