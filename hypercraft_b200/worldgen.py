"""Host-side input definitions for the mesher: chunk load lists and deterministic synthetic chunks.

* ``load_radius_positions`` restates misc/load_radius.cpp:16-35 (the sphere of meshable chunks and its
  one-chunk dilation, the resident set of ChunkPool::Update, client/src/ChunkPool.cpp:9-33).
* ``make_block`` follows Block(id, variant, transform), block/include/block/Block.hpp:46-47.
* The synthetic generators are this repo's own integer-hash inputs; they are
  machine independent, unlike the reference terrain (FastNoise2 is built -ffast-math with runtime SIMD dispatch).
"""
from __future__ import annotations

import numpy as np

CHUNK = 32

# block ids, block/register/blocks:1-29
AIR, STONE, COBBLESTONE, DIRT, GRASS_BLOCK, GRASS, SAND, DEAD_BUSH, GRAVEL, GLASS, SNOW, BLUE_ICE, ICE, SANDSTONE, \
    WATER, FLOWING_WATER, LEAVES, LOG, PLANK, APPLE, CACTUS, VINE, RED_MUSHROOM, BROWN_MUSHROOM, GLOWSTONE = range(25)

# variant bits per id (Trait.hpp:58-60, Liquid.hpp:305-306): grass block/grass 2, leaves/log/plank 3, flowing water 4
VARIANT_BITS = {GRASS_BLOCK: 2, GRASS: 2, LEAVES: 3, LOG: 3, PLANK: 3, FLOWING_WATER: 4}


def make_block(block_id: int, variant: int = 0, transform: int = 0) -> int:
    """Block::GetData() for (id, variant, transform): id | (variant | transform << variant_bits) << 8."""
    return int(block_id) | ((int(variant) | (int(transform) << VARIANT_BITS.get(block_id, 0))) << 8)


def load_radius_positions(radius: int):
    """(meshed, resident): chunk positions p with |p|^2 <= r^2 in the iteration order of load_radius.cpp, and the
    same set dilated by one chunk in every direction (first-seen order, like the reference list)."""
    r = int(radius)
    ax = np.arange(-r, r + 1, dtype=np.int32)
    x, y, z = np.meshgrid(ax, ax, ax, indexing="ij")
    m = (x * x + y * y + z * z) <= r * r
    meshed = np.stack([x[m], y[m], z[m]], axis=1).astype(np.int16)
    span = 2 * r + 3
    occ = np.zeros((span, span, span), dtype=bool)
    mi = meshed.astype(np.int32) + (r + 1)
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dz in (-1, 0, 1):
                occ[mi[:, 0] + dx, mi[:, 1] + dy, mi[:, 2] + dz] = True
    res = np.argwhere(occ).astype(np.int32) - (r + 1)
    return meshed, res.astype(np.int16)


def neighbourhood(pos) -> np.ndarray:
    """The 27 positions around ``pos`` (any order)."""
    p = np.asarray(pos, dtype=np.int16)
    d = np.array([[dx, dy, dz] for dx in (-1, 0, 1) for dy in (-1, 0, 1) for dz in (-1, 0, 1)], dtype=np.int16)
    return p[None, :] + d


# ---------------------------------------------------------------------------------------------------------------
# deterministic integer hashing (splitmix64)
# ---------------------------------------------------------------------------------------------------------------
def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def voxel_hash(seed: int, pos) -> np.ndarray:
    """u64 hash per voxel of chunk ``pos`` in memory order (index x + (y*32 + z)*32)."""
    idx = np.arange(CHUNK ** 3, dtype=np.uint64)
    px, py, pz = (np.uint64(int(c) & 0xFFFF) for c in pos)
    with np.errstate(over="ignore"):
        key = (np.uint64(seed) * np.uint64(0x100000001B3)) ^ (px << np.uint64(48)) ^ (py << np.uint64(32)) ^ (pz << np.uint64(16))
        return _splitmix64(_splitmix64(idx ^ key) + np.uint64(seed))


def column_hash(seed: int, pos) -> np.ndarray:
    return voxel_hash(seed ^ 0x5EED, pos)[:CHUNK * CHUNK]


def _coords():
    idx = np.arange(CHUNK ** 3)
    x = idx % CHUNK
    z = (idx // CHUNK) % CHUNK
    y = idx // (CHUNK * CHUNK)
    return x, y, z


def heights_consistent(blocks: np.ndarray, passable_vertical: np.ndarray) -> np.ndarray:
    """Sunlight heights the way the reference maintains them for one isolated chunk: lowest y such that every voxel
    at or above it lets vertical light pass (ChunkFloodSunlightTask.cpp:41-57 restricted to a single chunk)."""
    b = blocks.reshape(CHUNK, CHUNK, CHUNK)  # [y][z][x]
    ok = passable_vertical[b]  # bool [y][z][x]
    blocked = ~ok
    # height = 1 + highest blocked y, 0 if none
    ys = np.arange(1, CHUNK + 1).reshape(CHUNK, 1, 1)
    h = np.max(np.where(blocked, ys, 0), axis=0)
    return h.astype(np.int8).reshape(-1)  # index x + z*32


def synth_chunk(kind: str, seed: int, pos) -> tuple[np.ndarray, np.ndarray]:
    """Deterministic synthetic chunk: returns (blocks u16[32768], heights i8[1024]).

    kinds: air, stone, checker (3D checkerboard stone/air), transparent (leaves/water/glass/ice/air checker mixes),
    mix (40% stone + glowstone, leaves, water, glass, cross plants, cactus, vines, logs, flowing water 0..8),
    caves (smooth-ish solid with holes), flow (mostly flowing water of random levels next to water/stone/air),
    rawmeta (mix with out-of-table meta values, incl. flowing water meta > 8).
    """
    x, y, z = _coords()
    h = voxel_hash(seed, pos)
    hc = column_hash(seed, pos)
    r = (h & np.uint64(0xFFFF)).astype(np.int64)  # 16 bits
    r2 = ((h >> np.uint64(16)) & np.uint64(0xFFFF)).astype(np.int64)
    r3 = ((h >> np.uint64(32)) & np.uint64(0xFFFF)).astype(np.int64)
    blocks = np.zeros(CHUNK ** 3, dtype=np.uint16)
    rnd_heights = (hc % np.uint64(33)).astype(np.int8)
    if kind == "air":
        heights = np.zeros(1024, dtype=np.int8)
    elif kind == "stone":
        blocks[:] = STONE
        heights = np.full(1024, 32, dtype=np.int8)
    elif kind == "checker":
        par = (x + y + z + int(pos[0]) + int(pos[1]) + int(pos[2])) & 1
        blocks[par == 1] = STONE
        heights = rnd_heights
    elif kind == "transparent":
        par = (x + y + z) & 1
        choice = r % 8
        table = np.array([make_block(LEAVES, 0), make_block(LEAVES, 1), make_block(LEAVES, 4), WATER, GLASS, ICE,
                          make_block(LEAVES, 2), AIR], dtype=np.uint16)
        blocks[:] = np.where(par == 1, table[choice], np.where(r2 % 5 == 0, table[r3 % 8], AIR))
        heights = rnd_heights
    elif kind in ("mix", "rawmeta"):
        p = r % 1000
        blocks[:] = AIR
        blocks[p < 400] = STONE
        sel = (p >= 400) & (p < 410)
        blocks[sel] = GLOWSTONE
        sel = (p >= 410) & (p < 470)
        blocks[sel] = (LEAVES | ((r2[sel] % 5) << 8)).astype(np.uint16)
        sel = (p >= 470) & (p < 520)
        blocks[sel] = WATER
        sel = (p >= 520) & (p < 550)
        blocks[sel] = GLASS
        sel = (p >= 550) & (p < 570)
        blocks[sel] = ICE
        sel = (p >= 570) & (p < 600)
        blocks[sel] = (GRASS | ((r2[sel] % 4) << 8)).astype(np.uint16)
        sel = (p >= 600) & (p < 615)
        blocks[sel] = np.array([DEAD_BUSH, APPLE, RED_MUSHROOM, BROWN_MUSHROOM], dtype=np.uint16)[r2[sel] % 4]
        sel = (p >= 615) & (p < 635)
        blocks[sel] = CACTUS
        sel = (p >= 635) & (p < 660)
        blocks[sel] = (VINE | ((r2[sel] % 6) << 8)).astype(np.uint16)
        sel = (p >= 660) & (p < 690)
        blocks[sel] = (LOG | (((r2[sel] % 5) | ((r3[sel] % 3) << 3)) << 8)).astype(np.uint16)
        sel = (p >= 690) & (p < 720)
        blocks[sel] = (FLOWING_WATER | ((r2[sel] % 9) << 8)).astype(np.uint16)
        sel = (p >= 720) & (p < 740)
        blocks[sel] = (GRASS_BLOCK | ((r2[sel] % 4) << 8)).astype(np.uint16)
        sel = (p >= 740) & (p < 750)
        blocks[sel] = (PLANK | ((r2[sel] % 5) << 8)).astype(np.uint16)
        if kind == "rawmeta":
            sel = r3 % 7 == 0
            blocks[sel] = (blocks[sel] & 0xFF) | ((r2[sel] % 256) << 8).astype(np.uint16)
            sel = r3 % 97 == 1
            blocks[sel] = (25 + (r2[sel] % 231)).astype(np.uint16)  # unregistered ids ("None", Trait.hpp:24-34)
        heights = rnd_heights
    elif kind == "caves":
        # blocky value noise at 4-voxel cells -> large merged faces with holes
        cx, cy, cz = x // 4, y // 4, z // 4
        cell = (cx + 8 * (cz + 8 * cy)).astype(np.uint64)
        ch = _splitmix64(cell ^ np.uint64(seed * 7919 + (int(pos[0]) & 0xFF) * 31 + (int(pos[1]) & 0xFF) * 17 + (int(pos[2]) & 0xFF)))
        solid = (ch % np.uint64(100)) < np.uint64(55)
        blocks[solid] = np.where((ch[solid] >> np.uint64(8)) % np.uint64(3) == 0, DIRT, STONE).astype(np.uint16)
        sel = (~solid) & (r % 200 == 0)
        blocks[sel] = GLOWSTONE
        heights = None
    elif kind == "flow":
        p = r % 100
        blocks[:] = AIR
        blocks[p < 45] = (FLOWING_WATER | ((r2[p < 45] % 9) << 8)).astype(np.uint16)
        blocks[(p >= 45) & (p < 60)] = WATER
        blocks[(p >= 60) & (p < 75)] = STONE
        blocks[(p >= 75) & (p < 80)] = GLASS
        heights = rnd_heights
    else:
        raise ValueError(kind)
    if heights is None:
        heights = heights_consistent(blocks, _vertical_pass_table())
    return blocks, np.ascontiguousarray(heights, dtype=np.int8)


_VPASS = None


def _vertical_pass_table() -> np.ndarray:
    """bool[65536]: Block::GetVerticalLightPass (transparency == 2) from the committed registry blob."""
    global _VPASS
    if _VPASS is None:
        from . import registry

        reg = registry.load()
        _VPASS = reg.props["transparency"][reg.prop_index] == 2
    return _VPASS


# ---------------------------------------------------------------------------------------------------------------
# fallback terrain (only used by bench.py when the reference terrain generator, oracle/_ref, is not available)
# ---------------------------------------------------------------------------------------------------------------
def _lattice(seed: int, ix: np.ndarray, iz: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        k = (ix.astype(np.int64) & 0xFFFFF).astype(np.uint64) | ((iz.astype(np.int64) & 0xFFFFF).astype(np.uint64) << np.uint64(20))
        return (_splitmix64(k ^ np.uint64(seed * 0x9E3779B1)) & np.uint64(0xFFFF)).astype(np.int64)


def _value_noise(seed: int, wx: np.ndarray, wz: np.ndarray, cell: int) -> np.ndarray:
    """integer bilinear value noise in [0, 65535]"""
    ix, iz = np.floor_divide(wx, cell), np.floor_divide(wz, cell)
    fx, fz = wx - ix * cell, wz - iz * cell
    a, b = _lattice(seed, ix, iz), _lattice(seed, ix + 1, iz)
    c, d = _lattice(seed, ix, iz + 1), _lattice(seed, ix + 1, iz + 1)
    top = a * (cell - fx) + b * fx
    bot = c * (cell - fx) + d * fx
    return (top * (cell - fz) + bot * fz) // (cell * cell)


def synth_terrain_chunk(seed: int, pos) -> tuple[np.ndarray, np.ndarray]:
    """Deterministic height-field terrain (stone/dirt/grass, oceans, plants, a few glowstones)."""
    cx, cy, cz = (int(c) for c in pos)
    lx = np.arange(CHUNK)
    wx, wz = np.meshgrid(cx * CHUNK + lx, cz * CHUNK + lx, indexing="xy")  # [z][x]
    h = (_value_noise(seed, wx, wz, 64) - 32768) // 400 + (_value_noise(seed + 1, wx, wz, 16) - 32768) // 2600  # about +-95
    wy = (cy * CHUNK + np.arange(CHUNK)).reshape(CHUNK, 1, 1)  # [y][z][x]
    hh = h[None, :, :]
    blocks = np.zeros((CHUNK, CHUNK, CHUNK), dtype=np.uint16)
    blocks[wy < hh - 3] = STONE
    blocks[(wy >= hh - 3) & (wy < hh)] = DIRT
    top = np.broadcast_to(wy == hh, blocks.shape)
    blocks[top] = np.broadcast_to(np.where(hh > 1, GRASS_BLOCK, SAND), blocks.shape)[top].astype(np.uint16)
    blocks[(wy > hh) & (wy <= 0)] = WATER
    r = voxel_hash(seed, pos).reshape(CHUNK, CHUNK, CHUNK)
    plant = np.broadcast_to((wy == hh + 1) & (hh > 1), blocks.shape) & ((r % np.uint64(23)) == 0)
    blocks[plant] = GRASS
    lamp = np.broadcast_to(wy == hh + 1, blocks.shape) & ((r % np.uint64(4099)) == 1) & np.broadcast_to(hh > 1, blocks.shape)
    blocks[lamp] = GLOWSTONE
    surface = np.maximum(h, np.where(h < 0, 0, h))  # water is not vertical-light-pass: the ocean surface shades
    heights = np.clip(surface + 1 - cy * CHUNK, 0, CHUNK).astype(np.int8)
    lampcol = lamp.any(axis=0)
    if lampcol.any():
        lamp_top = np.max(np.where(lamp, np.arange(1, CHUNK + 1).reshape(CHUNK, 1, 1), 0), axis=0)
        heights = np.where(lampcol, np.maximum(heights, lamp_top), heights).astype(np.int8)
    return blocks.reshape(-1), heights.reshape(-1)
