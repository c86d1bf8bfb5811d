import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import hcoracle
from hypercraft_b200 import mesher as M, worldgen
R = int(os.environ.get("R", "20"))
meshed, resident = worldgen.load_radius_positions(R)
ref = hcoracle.RefWorld(); ref.generate(resident)
pb = M.PinnedArray((len(resident), 32768), np.uint16); ph = M.PinnedArray((len(resident), 1024), np.int8)
ref.get_chunks(resident, pb.array, ph.array)
ext = 2*R+3
m = M.Mesher(origin=(-R-1,)*3, extent=(ext,)*3, max_chunks=len(resident), max_tasks=len(meshed))
for it in range(4):
    t0 = time.perf_counter(); m.upload(resident, pb.array, ph.array); t1 = time.perf_counter(); m.sync(); t2 = time.perf_counter()
    m.mesh_async(meshed); t3 = time.perf_counter(); c = m.wait(); t4 = time.perf_counter(); r = m.fetch(); t5 = time.perf_counter()
    s = m.stats()
    print(f"upload call {1e3*(t1-t0):.1f} ms (+sync {1e3*(t2-t1):.1f}) | mesh submit {1e3*(t3-t2):.1f} wait {1e3*(t4-t3):.1f} | fetch {1e3*(t5-t4):.1f} | ev: copy {s['upload_copy_ms']:.1f} derive {s['derive_ms']:.2f} mesh {s['mesh_ms']:.2f} fetchcopy {s['fetch_copy_ms']:.1f}", flush=True)
