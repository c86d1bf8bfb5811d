"""Multi-GPU partitioning of a mesh-task list.

Chunks are independent given their 26 neighbours, so G GPUs need no exchange step: the task list is cut into G
contiguous x-slabs of (nearly) equal task count, and each GPU's resident set is its slab dilated by one chunk
(boundary chunk planes are simply uploaded to both neighbours).  No collective is involved on the data path."""
from __future__ import annotations

import numpy as np


def slab_bounds(positions: np.ndarray, world: int, weights=None) -> list[tuple[int, int]]:
    """Per rank [x_lo, x_hi) such that the slabs tile the x-range and hold balanced task counts -- or, with ``weights``
    (one per position), balanced total weight: most chunks of a load sphere are all air or all stone and cost next to
    nothing, so SURVEY 8e asks for a balance by estimated surface chunks."""
    pos = np.asarray(positions, dtype=np.int64).reshape(-1, 3)
    if len(pos) == 0:
        return [(0, 0)] * world
    w = np.ones(len(pos)) if weights is None else np.asarray(weights, dtype=np.float64).reshape(-1)
    x_min, x_max = int(pos[:, 0].min()), int(pos[:, 0].max()) + 1
    plane = np.bincount(pos[:, 0] - x_min, weights=w, minlength=x_max - x_min)  # weight of each x-plane
    before = np.concatenate([[0.0], np.cumsum(plane)])                            # weight left of boundary x_min + i
    cuts = [x_min]
    for r in range(1, world):
        i = int(np.argmin(np.abs(before - before[-1] * r / world)))               # the plane boundary nearest to the target
        cuts.append(max(x_min + i, cuts[-1]))
    cuts.append(x_max)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def surface_weights(positions: np.ndarray, band: int = 3, factor: float = 8.0) -> np.ndarray:
    """Cost estimate per task for terrain worlds: chunks within ``band`` chunk layers of y = 0 hold the surface and go
    through the full light + mesh kernel (about ``factor`` times the cost of a chunk that is classified away)."""
    pos = np.asarray(positions, dtype=np.int64).reshape(-1, 3)
    return 1.0 + factor * (np.abs(pos[:, 1]) <= band)


def shard_tasks(positions: np.ndarray, world: int, rank: int, weights=None) -> np.ndarray:
    pos = np.asarray(positions, dtype=np.int16).reshape(-1, 3)
    lo, hi = slab_bounds(pos, world, weights)[rank]
    sel = (pos[:, 0] >= lo) & (pos[:, 0] < hi)
    return pos[sel]


def resident_for(tasks: np.ndarray) -> np.ndarray:
    """The one-chunk dilation of a task set: every position a task reads."""
    t = np.asarray(tasks, dtype=np.int32).reshape(-1, 3)
    if len(t) == 0:
        return np.zeros((0, 3), dtype=np.int16)
    lo = t.min(axis=0) - 1
    span = t.max(axis=0) - lo + 2
    occ = np.zeros(tuple(int(s) for s in span), dtype=bool)
    ti = t - lo
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dz in (-1, 0, 1):
                occ[ti[:, 0] + dx, ti[:, 1] + dy, ti[:, 2] + dz] = True
    return (np.argwhere(occ) + lo).astype(np.int16)


def window_for(resident: np.ndarray) -> tuple[tuple[int, int, int], tuple[int, int, int]]:
    """(origin, extent) of the smallest hcm world window holding `resident`."""
    r = np.asarray(resident, dtype=np.int32).reshape(-1, 3)
    lo, hi = r.min(axis=0), r.max(axis=0)
    return tuple(int(v) for v in lo), tuple(int(v) for v in (hi - lo + 1))


# ---- edits on a sharded world (last sentence; BASELINE config 5 on N GPUs) -----------------------
# An edit belongs to the rank whose slab holds the edited chunk.  With resident_for_slab() that rank has the chunk's whole
# column of the world, i.e. exactly what the single-GPU world gives the sunlight chain (the chain walks along y, a slab is
# cut along x), so its hcm_set_blocks result -- changed heights and remesh fan-out -- is the undivided world's.  A rank that holds a touched chunk only as halo does NOT run the chain (the chunk above may be outside its
# dilation): it takes the owner's chunk as data (hcm_upload_chunks of that one chunk) before it meshes.  Remesh tasks go
# to the rank that owns their position; the owner's ctx reports those inside its slab, foreign_fanout() the rest.  All of it is host-side bookkeeping of a few bytes per edit; no collective.

def resident_for_slab(world_resident: np.ndarray, bounds, rank: int) -> np.ndarray:
    """Residency of a rank that also takes edits: every chunk of the undivided world's resident set in the x-planes
    [x_lo - 1, x_hi] -- a superset of resident_for(its tasks) by the fringe chunks that only a neighbouring slab's taller
    columns bring in.  With it the owner of a chunk holds that chunk's whole column of the world, so the sunlight chain
    stops on it exactly where it stops on one GPU, and "holds chunk c as halo" is a statement about c's x alone."""
    r = np.asarray(world_resident, dtype=np.int16).reshape(-1, 3)
    lo, hi = bounds[rank]
    if hi <= lo:
        return r[:0]
    return r[(r[:, 0] >= lo - 1) & (r[:, 0] <= hi)]


def owner_rank(chunk_pos, bounds) -> int:
    """Rank whose [x_lo, x_hi) holds the chunk (positions left / right of all slabs go to the first / last rank)."""
    x = int(chunk_pos[0])
    for r, (lo, hi) in enumerate(bounds):
        if x < hi and (x >= lo or r == 0):
            return r
    return len(bounds) - 1


def halo_ranks(chunk_pos, bounds) -> list[int]:
    """Ranks other than the owner whose slab, dilated by one chunk, reaches the chunk's x-plane."""
    x, own = int(chunk_pos[0]), owner_rank(chunk_pos, bounds)
    return [r for r, (lo, hi) in enumerate(bounds) if r != own and hi > lo and lo - 1 <= x <= hi]


def route_edits(edit_list, bounds) -> list[list]:
    """edit_list = [((cx,cy,cz), index, block), ...] -> per rank, the edits it applies with hcm_set_blocks (owner only)."""
    out = [[] for _ in bounds]
    for e in edit_list:
        out[owner_rank(e[0], bounds)].append(e)
    return out


def mirror_plan(touched_chunks, bounds) -> list[list[tuple[int, int, int]]]:
    """touched_chunks = chunk positions whose blocks or sunlight heights an owner changed -> per rank, the chunks it holds
    as halo and must take from the owner (re-upload) before its next mesh batch."""
    out = [[] for _ in bounds]
    for c in sorted({tuple(int(v) for v in c) for c in touched_chunks}):
        for r in halo_ranks(c, bounds):
            out[r].append(c)
    return out


def chunk_should_remesh(rel) -> bool:
    """ChunkShouldRemesh (client/include/client/ChunkMeshTask.inl:9-15): a block at `rel` (relative to a chunk's origin) can change that
    chunk's mesh iff its L1 distance to the box [-1,33]^3 is below 15 -- hcm_host::ShardedMeshDispatcher::ShouldRemesh in Python."""
    dist = 0
    for c in rel:
        c = int(c)
        dist += -1 - c if c < -1 else (c - 33 if c >= 33 else 0)
    return dist < 15


def remesh_offsets(block_pos) -> list[tuple[int, int, int]]:
    """Chunk offsets whose mesh task the reference pushes after an edit at chunk-local `block_pos` (ChunkSetBlockTask.cpp:100-107)."""
    return [(0, 0, 0)] + [(dx, dy, dz) for dx in (-1, 0, 1) for dy in (-1, 0, 1) for dz in (-1, 0, 1) if (dx, dy, dz) != (0, 0, 0)
                          and chunk_should_remesh((block_pos[0] - 32 * dx, block_pos[1] - 32 * dy, block_pos[2] - 32 * dz))]


def foreign_fanout(edit_list, sunlight_chunks, bounds, rank: int) -> set:
    """Remesh positions in OTHER ranks' slabs caused by `rank`'s edits.  Its hcm_set_blocks cannot report them -- one of their
    27 chunks lies outside its window, so the device counts them as pending -- hence they follow from the reference's rules on
    the host: per edit the neighbours passing ChunkShouldRemesh (ChunkSetBlockTask.cpp:100-107; every edit counts as a change,
    a superset by the no-op edits), per chunk with a changed sunlight height all 26 neighbours (ChunkSetSunlightTask.cpp:66-78).
    Same rule as hcm_host::ShardedMeshDispatcher::SetBlocks."""
    out = set()
    for cpos, index, _ in edit_list:
        bp = (int(index) & 31, int(index) >> 10, (int(index) >> 5) & 31)
        for d in remesh_offsets(bp):
            q = (int(cpos[0]) + d[0], int(cpos[1]) + d[1], int(cpos[2]) + d[2])
            if owner_rank(q, bounds) != rank:
                out.add(q)
    for c in sunlight_chunks:
        for dx in (-1, 1):
            for dy in (-1, 0, 1):
                for dz in (-1, 0, 1):
                    q = (int(c[0]) + dx, int(c[1]) + dy, int(c[2]) + dz)
                    if owner_rank(q, bounds) != rank:
                        out.add(q)
    return out


def split_tasks(positions, bounds) -> list[list[tuple[int, int, int]]]:
    """Remesh positions reported by the owners -> per rank, the ones it meshes (its own slab)."""
    out = [[] for _ in bounds]
    for p in sorted({tuple(int(v) for v in p) for p in positions}):
        out[owner_rank(p, bounds)].append(p)
    return out
