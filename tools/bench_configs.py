"""Secondary measurements for BASELINE.json configs 3 (worst-case face density) and 5 (edit-driven incremental remesh).
Not the headline benchmark (bench.py); prints one JSON object.  Run under gpurun."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import hcoracle
from hypercraft_b200 import edits, mesher as M, registry, worldgen
from util import synth_world

out = {}
# ---- config 3: every task a worst-case chunk; a 4x4x4 block of tasks inside a 6^3 window of the same kind ----------
cfg3 = {}
for kind in ["checker", "transparent", "mix", "flow", "caves"]:
    allpos = np.array([[x, y, z] for x in range(6) for y in range(6) for z in range(6)], dtype=np.int16)
    blocks = np.zeros((len(allpos), 32768), dtype=np.uint16); heights = np.zeros((len(allpos), 1024), dtype=np.int8)
    for i, p in enumerate(allpos):
        blocks[i], heights[i] = worldgen.synth_chunk(kind, 3, p)
    m = M.Mesher(origin=(0, 0, 0), extent=(6, 6, 6))
    m.upload(allpos, blocks, heights)
    tasks = np.array([[x, y, z] for x in range(1, 5) for y in range(1, 5) for z in range(1, 5)], dtype=np.int16)
    for _ in range(3):
        m.mesh_async(tasks); c = m.wait()
    ms = []
    for _ in range(5):
        m.mesh_async(tasks); c = m.wait(); ms.append(m.stats()["mesh_ms"])
    t = float(np.median(ms))
    cfg3[kind] = {"tasks": len(tasks), "mesh_ms": t, "chunks_per_s": len(tasks) / t * 1e3, "quads_per_s": c["n_vertices"] / 4 / t * 1e3,
                  "vertices_per_chunk": c["n_vertices"] / len(tasks), "algorithmic_GBps": (len(tasks) * 34 ** 3 * 3 + 8 * c["n_vertices"]) / t / 1e6}
    m.close()
out["config3_worst_case"] = cfg3

# ---- config 5: random single-block edits on reference terrain (r = 6), batches of 1 / 8 / 64 / 512 edits -------------
R = 6
meshed, resident = worldgen.load_radius_positions(R)
have_ref = hcoracle.have_ref()
if have_ref:
    ref = hcoracle.RefWorld(); ref.generate(resident); blocks, heights = ref.get_chunks(resident)
else:
    blocks = np.zeros((len(resident), 32768), dtype=np.uint16); heights = np.zeros((len(resident), 1024), dtype=np.int8)
    for i, p in enumerate(resident):
        blocks[i], heights[i] = worldgen.synth_terrain_chunk(12314524, p)
index = {tuple(int(c) for c in p): i for i, p in enumerate(resident)}
reg = registry.load(); vpass = reg.props["transparency"][reg.prop_index] == 2
ext = 2 * R + 3
m = M.Mesher(origin=(-R - 1,) * 3, extent=(ext,) * 3, max_chunks=len(resident), max_tasks=len(meshed))
m.upload(resident, blocks, heights); m.mesh(meshed)
meshable = {tuple(int(c) for c in p) for p in meshed}
surface = [tuple(int(c) for c in p) for p in meshed if abs(int(p[1])) <= 1 and max(abs(int(p[0])), abs(int(p[2]))) <= R - 2]
rng = np.random.default_rng(5)
cfg5 = {}
for batch in (1, 8, 64, 512):
    lat, ntasks = [], []
    for rep in range(12 if batch <= 64 else 5):
        t0 = time.perf_counter()
        dirty, tasks = set(), set()
        for e in range(batch):
            cpos = surface[int(rng.integers(0, len(surface)))]
            i = index[cpos]
            bx, bz = int(rng.integers(0, 32)), int(rng.integers(0, 32))
            by = int(np.clip(int(heights[i][bx + bz * 32]) + int(rng.integers(-2, 2)), 0, 31))
            blocks[i][bx + (by * 32 + bz) * 32] = [0, 1, 24, 14][int(rng.integers(0, 4))]
            up = index.get((cpos[0], cpos[1] + 1, cpos[2]))
            heights[i][bx + bz * 32] = edits.column_sunlight_height(blocks[i], bx, bz, int(heights[up][bx + bz * 32]) if up is not None else 0, vpass)
            dirty.add(i)
            for d in edits.remesh_offsets((bx, by, bz)):
                q = (cpos[0] + d[0], cpos[1] + d[1], cpos[2] + d[2])
                if q in meshable:
                    tasks.add(q)
        di = np.array(sorted(dirty))
        m.upload(resident[di], blocks[di], heights[di])
        res = m.mesh(np.array(sorted(tasks), dtype=np.int16))
        lat.append((time.perf_counter() - t0) * 1e3); ntasks.append(len(tasks))
    lat = np.array(lat[1:])
    cfg5[f"batch_{batch}"] = {"edits": batch, "remeshed_chunks_mean": float(np.mean(ntasks)), "latency_ms_p50": float(np.percentile(lat, 50)),
                              "latency_ms_p99": float(np.percentile(lat, 99)), "remeshed_chunks_per_s": float(np.mean(ntasks[1:]) / np.mean(lat) * 1e3),
                              "includes": "host edit bookkeeping (python), upload of dirty chunks, mesh kernel, fetch of the meshes"}
out["config5_edits"] = cfg5

# ---- config 5 through hcm_set_blocks: the edit, the sunlight chain and the fan-out happen on the device-resident world ----
cfg5d = {}
for batch in (1, 8, 64, 512):
    lat, ntasks, nsun = [], [], []
    for rep in range(12 if batch <= 64 else 5):
        batch_edits = np.zeros(batch, dtype=M.BLOCK_EDIT_DTYPE)
        for e in range(batch):
            cpos = surface[int(rng.integers(0, len(surface)))]
            i = index[cpos]
            bx, bz = int(rng.integers(0, 32)), int(rng.integers(0, 32))
            by = int(np.clip(int(heights[i][bx + bz * 32]) + int(rng.integers(-2, 2)), 0, 31))
            batch_edits[e]["chunk"], batch_edits[e]["index"], batch_edits[e]["block"] = cpos, bx + (by * 32 + bz) * 32, [0, 1, 24, 14][int(rng.integers(0, 4))]
        t0 = time.perf_counter()
        r = m.set_blocks(batch_edits)
        res = m.mesh(r["remesh"])
        lat.append((time.perf_counter() - t0) * 1e3); ntasks.append(len(r["remesh"])); nsun.append(len(r["sunlight"]))
    lat = np.array(lat[1:])
    cfg5d[f"batch_{batch}"] = {"edits": batch, "remeshed_chunks_mean": float(np.mean(ntasks)), "sunlight_changes_mean": float(np.mean(nsun)),
                               "latency_ms_p50": float(np.percentile(lat, 50)), "latency_ms_p99": float(np.percentile(lat, 99)),
                               "remeshed_chunks_per_s": float(np.mean(ntasks[1:]) / np.mean(lat) * 1e3),
                               "includes": "hcm_set_blocks (edit + sunlight chain kernels, host fan-out, derive of touched chunks), mesh kernel, fetch of the meshes; "
                                           "the remesh set is the reference's (27 chunks around every changed sunlight height)"}
out["config5_edits_on_device"] = cfg5d
print(json.dumps(out))
