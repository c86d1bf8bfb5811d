"""Developer tool: hcm_set_blocks + mesh kernel on the CPU emulation against the UNMODIFIED reference edit chain run live
(hcoracle.RefEditWorld: kSetBlock / kFloodSunlight / kSetSunlight / kMesh tasks under the reference's ChunkTaskPool::Run), on random
worlds and random edit batches.  Needs oracle/_ref.  Usage: python tools/sim_fuzz_edits.py [first_seed] [n_seeds]"""
import os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
sim = os.path.join(ROOT, "tests", "cusim")
subprocess.run(["make", "-C", sim], check=True, stdout=subprocess.DEVNULL)
import edit_oracle, hcoracle
from hypercraft_b200 import mesher as M
M.LIB_PATH = os.path.join(sim, "_build", "libhcm_sim.so")
first, n = int(sys.argv[1]) if len(sys.argv) > 1 else 100, int(sys.argv[2]) if len(sys.argv) > 2 else 5
bad = batches = 0
t0 = time.time()
for seed in range(first, first + n):
    allpos, chunks = edit_oracle.edit_test_world(seed)
    rw = hcoracle.RefEditWorld(4)
    for p in allpos:
        rw.set_chunk(p, *chunks[tuple(int(c) for c in p)])
    m = M.Mesher(origin=(-2, -2, -2), extent=(5, 5, 5))
    m.upload(allpos, np.stack([chunks[tuple(int(c) for c in p)][0] for p in allpos]), np.stack([chunks[tuple(int(c) for c in p)][1] for p in allpos]))
    keys = sorted(chunks)
    for name, batch in edit_oracle.edit_test_batches(chunks, seed=seed + 7, n_random=5):
        before = {k: rw.get_chunk(k) for k in keys}
        pushed, high, still, meshes = rw.apply(batch, with_meshes=True)
        after = {k: rw.get_chunk(k) for k in keys}
        want_sun = sorted([list(k), int(i), int(before[k][1][i]), int(after[k][1][i])] for k in keys for i in np.nonzero(before[k][1] != after[k][1])[0])
        want_changed = int(sum(int((before[k][0] != after[k][0]).sum()) for k in keys))
        want = {}
        for i, p in enumerate(pushed):
            e = want.setdefault(tuple(int(v) for v in p), {"high": 0})
            e["high"] |= int(high[i])
            e["mesh"] = meshes[i]
        r = m.set_blocks(batch)
        got_sun = sorted([[int(v) for v in s["chunk"]], int(s["index"]), int(s["old_height"]), int(s["new_height"])] for s in r["sunlight"])
        got = {tuple(int(v) for v in p): int(f) for p, f in zip(r["remesh"], r["remesh_flags"])}
        ok = r["n_changed_blocks"] == want_changed and got_sun == want_sun and set(got) == set(want) and all(bool(got[p] & 1) == bool(want[p]["high"]) for p in want)
        if ok and len(r["remesh"]):
            res = m.mesh(r["remesh"])
            for t, p in enumerate(r["remesh"]):
                g, w = res.serialise(t), want[tuple(int(v) for v in p)]["mesh"]
                ok = ok and g.size == w.size and bool((g == w).all())
        batches += 1
        if not ok:
            bad += 1
            print(f"MISMATCH world {seed} batch {name}: changed {r['n_changed_blocks']}/{want_changed} sun {len(got_sun)}/{len(want_sun)} remesh {len(got)}/{len(want)}", flush=True)
        for k in keys:  # the batch generator reads the current heights
            chunks[k][0][:] = after[k][0]
            chunks[k][1][:] = after[k][1]
    m.close()
print(f"{batches} batches, {bad} mismatches, {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
