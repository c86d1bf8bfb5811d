"""Regenerates tests/golden/synth_golden.json from the UNMODIFIED reference (oracle/_ref/libhcref.so).
Run where /root/reference exists (`make -C oracle ref` first).  The inputs are this repo's integer-hash synthetic
chunks (hypercraft_b200.worldgen.synth_chunk), hence machine independent; the recorded values are SHA-256 digests of
the reference's serialised meshes (layout: oracle/ref_shim.cpp) and of its [-1,32]^3 light bytes."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import hcoracle  # noqa: E402
from util import synth_world  # noqa: E402

CASES = [(k, s) for k in ["air", "stone", "checker", "transparent", "mix", "rawmeta", "caves", "flow"] for s in (1, 2, 3)]
MIXED = [("stone", {(1, 0, 0): "air", (0, 1, 0): "mix", (0, 0, -1): "flow"}),
         ("air", {(0, -1, 0): "stone", (1, 0, 0): "mix"}),
         ("flow", {(1, 0, 0): "mix", (-1, 0, 0): "flow", (0, 0, 1): "stone", (0, 0, -1): "rawmeta"}),
         ("caves", {(0, 1, 0): "air", (0, -1, 0): "stone", (1, 1, 0): "mix"})]


# height-field terrain from the repo's integer hash (worldgen.synth_terrain_chunk): (seed, centre chunk), surface-rich ones
TERRAIN = [(12314524, (0, 0, 0)), (12314524, (0, 1, 2)), (12314524, (2, -1, 1)), (7, (-2, 0, -1)), (7, (0, 0, 1)), (7, (-1, -1, -2))]


def record(ref, pos, blocks, heights, centre=(0, 0, 0)):
    for i, p in enumerate(pos):
        ref.set_chunk(p, blocks[i], heights[i])
    buf = ref.mesh(centre)
    ms = hcoracle.parse_meshes(buf)
    return dict(mesh_sha256=hashlib.sha256(buf.tobytes()).hexdigest(), light_sha256=hashlib.sha256(ref.light34(centre).tobytes()).hexdigest(),
                meshes=[[m["transparent"], int(len(m["vertices"])), int(len(m["indices"]))] for m in ms])


def terrain_world(seed, centre):
    from hypercraft_b200 import worldgen

    pos = worldgen.neighbourhood(centre)
    chunks = [worldgen.synth_terrain_chunk(seed, p) for p in pos]
    return pos, [c[0] for c in chunks], [c[1] for c in chunks]


if __name__ == "__main__":
    out = {"source": "unmodified reference ChunkTaskRunner<kMesh>::Run via oracle/_ref", "cases": {}, "mixed": []}
    for kind, seed in CASES:
        pos, blocks, heights, _ = synth_world(kind, seed)
        out["cases"][f"{kind}/{seed}"] = record(hcoracle.RefWorld(), pos, blocks, heights)
    for i, (kind, by_off) in enumerate(MIXED):
        pos, blocks, heights, _ = synth_world(kind, 10 + i, kinds_by_offset=by_off)
        out["mixed"].append(record(hcoracle.RefWorld(), pos, blocks, heights))
    out["terrain"] = [record(hcoracle.RefWorld(), *terrain_world(seed, centre), centre=centre) for seed, centre in TERRAIN]
    json.dump(out, open(os.path.join(HERE, "synth_golden.json"), "w"), indent=1)
    print("wrote synth_golden.json with", len(out["cases"]) + len(out["mixed"]) + len(out["terrain"]), "cases")
