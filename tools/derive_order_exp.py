import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import hcoracle
from hypercraft_b200 import mesher as M, worldgen
R = 20
meshed, resident = worldgen.load_radius_positions(R)
ref = hcoracle.RefWorld(); ref.generate(resident)
for name, order in [("x-major (as listed)", np.arange(len(resident))), ("cx fastest", np.lexsort((resident[:,0], resident[:,2], resident[:,1]))), ("random", np.random.default_rng(0).permutation(len(resident)))]:
    res = resident[order]
    pb = M.PinnedArray((len(res), 32768), np.uint16); ph = M.PinnedArray((len(res), 1024), np.int8)
    ref.get_chunks(res, pb.array, ph.array)
    ext = 2*R+3
    m = M.Mesher(origin=(-R-1,)*3, extent=(ext,)*3, max_chunks=len(res), max_tasks=len(meshed))
    m.upload(res, pb.array, ph.array); m.sync()
    ts=[]
    for _ in range(5):
        m.rederive(); m.sync(); ts.append(m.stats()["derive_ms"])
    m.mesh_async(meshed); c = m.wait(); s = m.stats()
    print(f"{name:22s} derive_ms {np.min(ts):.3f} (runs {[round(t,3) for t in ts]}) mesh_ms {s['mesh_ms']:.3f} verts {c['n_vertices']}", flush=True)
    m.close(); pb.free(); ph.free()
