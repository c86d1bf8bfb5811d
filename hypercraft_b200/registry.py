"""Registry blob (``hcm_registry`` of include/hcm.h): the flat mirror of the reference's compile-time block and
texture tables (block/include/block/Block.hpp:35-41, BlockProperty.hpp:17-29, BlockMesh.hpp:15-49,
texture/internal/generated/block_texture_transparency.inl:1-3).

``data/registry.bin`` is produced by oracle/export_registry.py from the reference's own registry code; a HyperCraft
build would export it at start-up with the snippet in INTEGRATION.md."""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

BLOB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "registry.bin")
MAGIC = 0x4752434D

PROP_DTYPE = np.dtype([("tex", np.uint16, 6), ("transparency", np.uint8), ("light_level", np.uint8), ("collision", np.uint8),
                       ("custom_mesh", np.uint8), ("liquid_role", np.uint8), ("liquid_obstacle", np.uint8), ("_pad", np.uint8, 2)])
FACE_DTYPE = np.dtype([("axis", np.uint8), ("light_face", np.uint8), ("render_face", np.uint8), ("_pad0", np.uint8),
                       ("texture", np.uint16), ("_pad1", np.uint16), ("verts", np.uint8, (4, 4))])
MESH_DTYPE = np.dtype([("face_count", np.uint32), ("dynamic_kind", np.uint32), ("liquid_source_id", np.uint8),
                       ("liquid_flow_id", np.uint8), ("liquid_height_unit", np.uint8), ("liquid_variant", np.uint8),
                       ("liquid_flow_tex", np.uint16), ("liquid_obstacle_rule", np.uint8), ("_pad", np.uint8),
                       ("faces", FACE_DTYPE, 16)])
assert PROP_DTYPE.itemsize == 20 and FACE_DTYPE.itemsize == 24 and MESH_DTYPE.itemsize == 400


@dataclass
class Registry:
    blob: np.ndarray
    prop_index: np.ndarray  # u8[65536]
    props: np.ndarray  # PROP_DTYPE[n_props]
    meshes: np.ndarray  # MESH_DTYPE[n_meshes]
    tex_flags: np.ndarray  # u8[n_textures]


def parse(blob: np.ndarray) -> Registry:
    blob = np.ascontiguousarray(blob, dtype=np.uint8)
    hdr = blob[:32].view(np.uint32)
    if hdr[0] != MAGIC or hdr[1] != 1:
        raise ValueError("not an hcm registry blob")
    n_props, n_meshes, n_tex = int(hdr[2]), int(hdr[3]), int(hdr[4])
    off = 32
    prop_index = blob[off:off + 65536]
    off += 65536
    props = blob[off:off + 20 * n_props].view(PROP_DTYPE)
    off += 20 * n_props
    meshes = blob[off:off + 400 * n_meshes].view(MESH_DTYPE)
    off += 400 * n_meshes
    tex_flags = blob[off:off + n_tex]
    if off + n_tex > blob.size:
        raise ValueError("truncated registry blob")
    return Registry(blob, prop_index, props, meshes, tex_flags)


def load(path: str = BLOB_PATH) -> Registry:
    return parse(np.fromfile(path, dtype=np.uint8))
