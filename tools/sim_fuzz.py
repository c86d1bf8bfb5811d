"""Developer tool: differential fuzzing of the kernels on the CPU emulation (tests/cusim) against the CPU oracle port -- random seeds
for every synthetic kind, random mixes of kinds over the 27 chunks, random terrain centres, light stage and mesh.  Usage:
    python tools/sim_fuzz.py [first_seed] [n_seeds]
Prints every mismatch; exit code 1 if there was one."""
import os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
sim = os.path.join(ROOT, "tests", "cusim")
subprocess.run(["make", "-C", sim], check=True, stdout=subprocess.DEVNULL)
import hcoracle
from hypercraft_b200 import mesher as M, registry, worldgen
from util import synth_world
M.LIB_PATH = os.path.join(sim, "_build", "libhcm_sim.so")
first, n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000, int(sys.argv[2]) if len(sys.argv) > 2 else 50
KINDS = ["air", "stone", "checker", "transparent", "mix", "rawmeta", "caves", "flow"]
port = hcoracle.PortOracle(registry.load().blob)
m = M.Mesher(origin=(-2, -2, -2), extent=(5, 5, 5))
bad = cases = 0
t0 = time.time()
for seed in range(first, first + n):
    rng = np.random.default_rng(seed)
    trials = [(k, None) for k in ("mix", "caves", "flow", "transparent", "rawmeta")]
    offs = [(x, y, z) for x in (-1, 0, 1) for y in (-1, 0, 1) for z in (-1, 0, 1)]
    for _ in range(3):  # random neighbourhood mixes
        trials.append((KINDS[int(rng.integers(len(KINDS)))], {o: KINDS[int(rng.integers(len(KINDS)))] for o in offs if rng.integers(2)}))
    for kind, by_off in trials:
        pos, blocks, heights, chunks = synth_world(kind, seed, kinds_by_offset=by_off)
        if rng.integers(2):  # random sunlight heights on some chunks (the reference only requires 0..32)
            i = int(rng.integers(27))
            heights[i] = rng.integers(0, 33, size=1024).astype(np.int8)
            chunks[tuple(int(c) for c in pos[i])] = (blocks[i], heights[i])
        m.upload(pos, blocks, heights)
        got = m.mesh(np.array([[0, 0, 0]], dtype=np.int16)).serialise(0)
        want = port.mesh(chunks, (0, 0, 0))
        lg, lw = m.debug_light((0, 0, 0)), port.light34(chunks, (0, 0, 0))
        cases += 1
        if got.size != want.size or not (got == want).all() or not (lg == lw).all():
            bad += 1
            print(f"MISMATCH seed {seed} kind {kind} mix {by_off}: mesh {got.size} vs {want.size} bytes, light diffs {(lg != lw).sum()}", flush=True)
    # terrain centre
    centre = tuple(int(v) for v in rng.integers(-3, 4, size=3))
    pos = worldgen.neighbourhood((0, 0, 0))
    tb = np.zeros((27, 32768), dtype=np.uint16); th = np.zeros((27, 1024), dtype=np.int8); tc = {}
    for i, p in enumerate(pos):
        tb[i], th[i] = worldgen.synth_terrain_chunk(seed, tuple(int(a + b) for a, b in zip(p, centre)))
        tc[tuple(int(c) for c in p)] = (tb[i], th[i])
    m.upload(pos, tb, th)
    got = m.mesh(np.array([[0, 0, 0]], dtype=np.int16)).serialise(0)
    want = port.mesh(tc, (0, 0, 0))
    cases += 1
    if got.size != want.size or not (got == want).all():
        bad += 1
        print(f"MISMATCH terrain seed {seed} centre {centre}", flush=True)
print(f"{cases} cases, {bad} mismatches, {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
