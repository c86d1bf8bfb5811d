"""Host logic of edit-driven remeshing (BASELINE.json config 5), restated from the reference callers of the mesh task.

* ``chunk_should_remesh``  -- ChunkShouldRemesh, client/include/client/ChunkMeshTask.inl:9-15: a block at position
  ``rel`` (relative to a chunk's origin) affects that chunk's mesh iff its L1 distance to the box [-1,33]^3 is < 15
  (IsBorderLightInterfere with level 15, client/include/client/BlockLightAlgo.hpp:32-42).
* ``remesh_offsets``       -- the fan-out of ChunkTaskRunner<kSetBlock>::Run, client/src/ChunkSetBlockTask.cpp:100-107:
  the edited chunk always, neighbour i iff ChunkShouldRemesh(block_pos - cmp_i * 32).
* ``column_sunlight_height`` -- ChunkTaskRunner<kFloodSunlight>::Run, client/src/ChunkFloodSunlightTask.cpp:41-57.
"""
from __future__ import annotations

import numpy as np

CHUNK = 32


def chunk_should_remesh(rel) -> bool:
    dist = 0
    for c in rel:
        c = int(c)
        if c < -1:
            dist += -1 - c
        elif c >= CHUNK + 1:
            dist += c - (CHUNK + 1)
    return 15 > dist


def remesh_offsets(block_pos) -> list[tuple[int, int, int]]:
    """Chunk offsets (dx,dy,dz) whose mesh task must be re-queued after an edit at ``block_pos`` (chunk-local 0..31)."""
    out = [(0, 0, 0)]
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dz in (-1, 0, 1):
                if (dx, dy, dz) == (0, 0, 0):
                    continue
                rel = (block_pos[0] - dx * CHUNK, block_pos[1] - dy * CHUNK, block_pos[2] - dz * CHUNK)
                if chunk_should_remesh(rel):
                    out.append((dx, dy, dz))
    return out


def column_sunlight_height(blocks: np.ndarray, x: int, z: int, up_height: int, vertical_pass: np.ndarray) -> int:
    """New sunlight height of column (x,z) of a chunk: 32 if the chunk above is not fully lit in that column, else one
    above the highest voxel that does not let vertical light pass (0 if none)."""
    if up_height > 0:
        return CHUNK
    col = blocks.reshape(CHUNK, CHUNK, CHUNK)[:, z, x]  # [y]
    y = CHUNK - 1
    while y >= 0 and vertical_pass[int(col[y])]:
        y -= 1
    return y + 1
