"""TEST INFRASTRUCTURE -- the checker of hcm_set_blocks (not part of the shipped product; only tests/ and tools/ import it).

Host logic of edit-driven remeshing (BASELINE.json config 5), restated from the reference callers of the mesh task and PINNED to
them: tests/test_oracle_golden.py runs the same edit batches through the reference's own kSetBlock / kFloodSunlight / kSetSunlight
tasks (oracle/_ref, hcoracle.RefEditWorld) and through this file, live where oracle/_ref exists and against
tests/golden/edit_chain_golden.json everywhere.

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


def apply_block_edits(chunks: dict, edit_list, vertical_pass: np.ndarray, with_flags: bool = False):
    """Host restatement of what ``hcm_set_blocks`` does on the device: the reference's kSetBlock -> kFloodSunlight ->
    kSetSunlight -> kFloodSunlight(below) chain (client/src/ChunkSetBlockTask.cpp:78-107, ChunkFloodSunlightTask.cpp:41-57,
    ChunkSetSunlightTask.cpp:42-78) run to its settled state on ``chunks`` = {(cx,cy,cz): (blocks u16[32768], heights i8[1024])},
    which is modified in place.  ``edit_list`` = [((cx,cy,cz), index, block), ...]; one pending update per voxel, the last
    one given wins.  Returns (n_changed_blocks, sunlight_changes, remesh_positions):
      sunlight_changes = {((cx,cy,cz), xz): (old, new)}; remesh_positions = set of chunk positions, NOT filtered by residency.
    ``with_flags``: a fourth value {position: 1 (pushed by a changed block: Push<kMesh>(pos, true), ChunkSetBlockTask.cpp:119) |
    2 (pushed by a changed sunlight height: Push<kMesh>(pos, false), ChunkSetSunlightTask.cpp:84)}.
    """
    pending = {}
    for cpos, index, block in edit_list:
        pending[(tuple(int(v) for v in cpos), int(index))] = int(block)
    changed = []
    for (cpos, index), block in pending.items():
        b, _ = chunks[cpos]
        if int(b[index]) == block:  # ChunkSetBlockTask.cpp:89-90
            continue
        b[index] = block
        changed.append((cpos, index))
    remesh = set()
    flags = {}
    for cpos, index in changed:
        bp = (index & 31, index >> 10, (index >> 5) & 31)
        for d in remesh_offsets(bp):
            q = (cpos[0] + d[0], cpos[1] + d[1], cpos[2] + d[2])
            remesh.add(q)
            flags[q] = flags.get(q, 0) | 1
    sun = {}
    # columns top-down: a chain settles every chunk it walks through
    for cpos, index in sorted(changed, key=lambda t: -t[0][1]):
        x, z = index & 31, (index >> 5) & 31
        xz = x + z * 32
        c = cpos
        while True:
            up = (c[0], c[1] + 1, c[2])
            if c not in chunks or up not in chunks:  # Pop needs both chunks, ChunkFloodSunlightTask.cpp:20-23
                break
            b, h = chunks[c]
            nh = column_sunlight_height(b, x, z, int(chunks[up][1][xz]), vertical_pass)
            old = int(h[xz])
            if nh == old:
                break
            h[xz] = nh
            key = (c, xz)
            sun[key] = (sun[key][0] if key in sun else old, nh)
            for dx in (-1, 0, 1):  # ChunkSetSunlightTask.cpp:66-78: every neighbour, the test never depends on i
                for dy in (-1, 0, 1):
                    for dz in (-1, 0, 1):
                        q = (c[0] + dx, c[1] + dy, c[2] + dz)
                        remesh.add(q)
                        flags[q] = flags.get(q, 0) | 2
            c = (c[0], c[1] - 1, c[2])
    sun = {k: v for k, v in sun.items() if v[0] != v[1]}
    if with_flags:
        return len(changed), sun, remesh, flags
    return len(changed), sun, remesh


def edit_test_world(seed: int = 78):
    """The 5 x 5 x 5 world of the edit-chain tests: integer-hash height-field terrain (machine independent)."""
    from hypercraft_b200 import worldgen

    allpos = np.array([[x, y, z] for x in range(-2, 3) for y in range(-2, 3) for z in range(-2, 3)], dtype=np.int16)
    chunks = {}
    for p in allpos:
        b, h = worldgen.synth_terrain_chunk(seed, p)
        chunks[tuple(int(c) for c in p)] = (b.copy(), h.copy())
    return allpos, chunks


def edit_test_batches(chunks: dict, seed: int = 21, n_random: int = 8):
    """Generator of the seeded edit batches shared by tests/golden/make_edit_golden.py and the tests: a shaft dug from the sky through two
    chunks in ONE batch (the chain has to run top-down), a lid that closes it again, then random batches around the terrain surface
    with duplicates, no-ops, several edits per world column and neighbours across chunk borders.  A batch depends on the heights the
    previous ones left behind, so the consumer must apply each batch to `chunks` before asking for the next."""
    from hypercraft_b200 import worldgen

    rng = np.random.default_rng(seed)
    palette = [0, 1, 24, 14, worldgen.make_block(worldgen.FLOWING_WATER, 5), 9, worldgen.make_block(worldgen.LEAVES, 2), 20, 5]
    x, z = 9, 20
    yield "shaft", [((0, cy, 0), x + (y * 32 + z) * 32, 0) for cy in (1, 0) for y in range(32)]
    yield "lid", [((0, 1, 0), x + (31 * 32 + z) * 32, 1)]
    for it in range(n_random):
        batch = []
        for _ in range(int(rng.integers(1, 40))):
            cpos = (int(rng.integers(-1, 2)), int(rng.integers(-1, 2)), int(rng.integers(-1, 2)))
            bx, bz = int(rng.integers(0, 32)), int(rng.integers(0, 32))
            h = int(chunks[cpos][1][bx + bz * 32])
            by = int(np.clip(h + int(rng.integers(-4, 3)), 0, 31)) if rng.integers(0, 3) else int(rng.integers(0, 32))
            batch.append((cpos, bx + (by * 32 + bz) * 32, palette[int(rng.integers(0, len(palette)))]))
            if rng.integers(0, 5) == 0:
                batch.append(batch[-1][:2] + (palette[int(rng.integers(0, len(palette)))],))  # same voxel twice: the last one wins
            if rng.integers(0, 4) == 0:  # another edit in the same world column, other chunk
                batch.append(((cpos[0], int(rng.integers(-1, 2)), cpos[2]), bx + (int(rng.integers(0, 32)) * 32 + bz) * 32, 0))
        yield f"random {it}", batch
