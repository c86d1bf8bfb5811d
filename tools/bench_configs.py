"""Secondary measurements for BASELINE.json configs[2] (worst-case face density) and configs[4] (edit-driven incremental remesh).
bench.py prints them as `secondary_configs` at N = 1; `python tools/bench_configs.py` prints them alone.  Every result is compared
with the CPU arm chunk for chunk (digest of the serialised meshes), like the headline's parity gate."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

H_BYTES = 34 ** 3 * 3


def _checker(chunks: dict, threads: int):
    """CPU arm over a dict of chunks: the unmodified reference when oracle/_ref is built, else the port."""
    import hcoracle
    from hypercraft_b200 import registry

    if hcoracle.have_ref() and not os.environ.get("HCM_BENCH_FORCE_PORT"):
        ref = hcoracle.RefWorld()
        for p, (b, h) in chunks.items():
            ref.set_chunk(p, b, h)
        return "reference", lambda pos: ref.mesh_batch(pos, threads=threads, per_chunk=True)[1]
    port = hcoracle.PortOracle(registry.load().blob)
    return "port", lambda pos: port.mesh_batch(chunks, pos, threads=threads)[1]


def config3(device: int = 0, threads: int = 8, window: int = 5, seed: int = 3) -> dict:
    """Every task a worst-case chunk: the (window-2)^3 inner tasks of a window^3 block of chunks of one synthetic kind
    (hypercraft_b200/worldgen.py: 3-D checkerboard, transparency mixes, the 40 % fill mix, flowing water, caves)."""
    import hcoracle
    from hypercraft_b200 import mesher as M
    from hypercraft_b200 import worldgen

    out = {}
    rng = range(window)
    allpos = np.array([[x, y, z] for x in rng for y in rng for z in rng], dtype=np.int16)
    tasks = np.array([[x, y, z] for x in range(1, window - 1) for y in range(1, window - 1) for z in range(1, window - 1)], dtype=np.int16)
    for kind in ["checker", "transparent", "mix", "flow", "caves"]:
        blocks = np.zeros((len(allpos), 32768), dtype=np.uint16)
        heights = np.zeros((len(allpos), 1024), dtype=np.int8)
        for i, p in enumerate(allpos):
            blocks[i], heights[i] = worldgen.synth_chunk(kind, seed, p)
        m = M.Mesher(origin=(0, 0, 0), extent=(window,) * 3, device=device)
        m.upload(allpos, blocks, heights)
        for _ in range(3):
            m.mesh_async(tasks)
            c = m.wait()
        ms = []
        for _ in range(5):
            m.mesh_async(tasks)
            c = m.wait()
            ms.append(m.stats()["mesh_ms"])
        t = float(np.median(ms))
        res = m.mesh(tasks)
        kind_name, check = _checker({tuple(int(v) for v in p): (blocks[i], heights[i]) for i, p in enumerate(allpos)}, threads)
        want, got = check(tasks), hcoracle.digest_batch(res, threads)
        out[kind] = {"tasks": len(tasks), "mesh_ms": t, "chunks_per_s": len(tasks) / t * 1e3, "quads_per_s": c["n_vertices"] / 4 / t * 1e3,
                     "vertices_per_chunk": c["n_vertices"] / len(tasks), "algorithmic_GBps": (len(tasks) * H_BYTES + 8 * c["n_vertices"]) / t / 1e6,
                     "parity": {"checker": kind_name, "byte_compared_chunks": len(tasks), "byte_mismatches": int((got[:, :4] != want[:, :4]).any(axis=1).sum())}}
        m.close()
    out["note"] = f"{len(tasks)} tasks per kind occupy {len(tasks)} of the 148 persistent CTAs: a latency figure per worst-case chunk, not a throughput one"
    return out


def config5_device(device: int = 0, threads: int = 8, radius: int = 6, seed: int = 5) -> dict:
    """Random single-block edits near the surface of reference terrain (r = 6), in batches of 1 / 8 / 64 / 512, through hcm_set_blocks: the
    edit, the sunlight chain and the remesh fan-out happen on the device-resident world; the timed region runs from the call to the
    fetched meshes.  Afterwards the host applies the same edits and the reported sunlight changes to its own chunks (what a client
    does with its Chunk objects) and the CPU arm meshes every chunk the edits ever touched: the device's meshes must be identical."""
    import hcoracle
    from hypercraft_b200 import mesher as M
    from hypercraft_b200 import registry, worldgen

    meshed, resident = worldgen.load_radius_positions(radius)
    if hcoracle.have_ref() and not os.environ.get("HCM_BENCH_FORCE_PORT"):
        gen = hcoracle.RefWorld()
        gen.generate(resident, threads=threads)
        blocks, heights = gen.get_chunks(resident)
        del gen
    else:
        blocks = np.zeros((len(resident), 32768), dtype=np.uint16)
        heights = np.zeros((len(resident), 1024), dtype=np.int8)
        for i, p in enumerate(resident):
            blocks[i], heights[i] = worldgen.synth_terrain_chunk(12314524, p)
    index = {tuple(int(c) for c in p): i for i, p in enumerate(resident)}
    ext = 2 * radius + 3
    m = M.Mesher(origin=(-radius - 1,) * 3, extent=(ext,) * 3, device=device, max_chunks=len(resident), max_tasks=len(meshed))
    m.upload(resident, blocks, heights)
    m.mesh(meshed)
    surface = [tuple(int(c) for c in p) for p in meshed if abs(int(p[1])) <= 1 and max(abs(int(p[0])), abs(int(p[2]))) <= radius - 2]
    rng = np.random.default_rng(seed)
    out, touched = {}, set()
    for batch in (1, 8, 64, 512):
        lat, ntasks, nsun = [], [], []
        for rep in range(12 if batch <= 64 else 5):
            e = np.zeros(batch, dtype=M.BLOCK_EDIT_DTYPE)
            for k in range(batch):
                cpos = surface[int(rng.integers(0, len(surface)))]
                i = index[cpos]
                bx, bz = int(rng.integers(0, 32)), int(rng.integers(0, 32))
                by = int(np.clip(int(heights[i][bx + bz * 32]) + int(rng.integers(-2, 2)), 0, 31))
                e[k]["chunk"], e[k]["index"], e[k]["block"] = cpos, bx + (by * 32 + bz) * 32, [0, 1, 24, 14][int(rng.integers(0, 4))]
            t0 = time.perf_counter()
            r = m.set_blocks(e)
            m.mesh(r["remesh"])
            lat.append((time.perf_counter() - t0) * 1e3)
            ntasks.append(len(r["remesh"])), nsun.append(len(r["sunlight"]))
            # the host's copy: one pending update per voxel, the last one wins (like the ctx); then the reported heights
            for k in range(batch):
                blocks[index[tuple(int(v) for v in e[k]["chunk"])]][int(e[k]["index"])] = int(e[k]["block"])
            for s in r["sunlight"]:
                heights[index[tuple(int(v) for v in s["chunk"])]][int(s["index"])] = int(s["new_height"])
            touched.update(tuple(int(v) for v in p) for p in r["remesh"])
        lat = np.array(lat[1:])
        out[f"batch_{batch}"] = {"edits": batch, "remeshed_chunks_mean": float(np.mean(ntasks)), "sunlight_changes_mean": float(np.mean(nsun)),
                                 "latency_ms_p50": float(np.percentile(lat, 50)), "latency_ms_p99": float(np.percentile(lat, 99)),
                                 "remeshed_chunks_per_s": float(np.mean(ntasks[1:]) / np.mean(lat) * 1e3)}
    # parity: every chunk the edits touched, device-resident world against the host's edited copy
    tp = np.array(sorted(touched), dtype=np.int16).reshape(-1, 3)
    res = m.mesh(tp)
    kind_name, check = _checker({tuple(int(v) for v in p): (blocks[i], heights[i]) for i, p in enumerate(resident)}, threads)
    want, got = check(tp), hcoracle.digest_batch(res, threads)
    out["parity"] = {"checker": kind_name, "byte_compared_chunks": int(len(tp)), "byte_mismatches": int((got[:, :4] != want[:, :4]).any(axis=1).sum())}
    out["includes"] = ("hcm_set_blocks (edit + sunlight chain kernels, host fan-out, derive of touched chunks), mesh kernel, fetch of the meshes; the remesh "
                       "set is the reference's (27 chunks around every changed sunlight height)")
    m.close()
    return out


if __name__ == "__main__":
    print(json.dumps({"config3_worst_case": config3(), "config5_edits_on_device": config5_device()}))
