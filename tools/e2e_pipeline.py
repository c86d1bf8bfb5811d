"""Timeline of hcm_upload_and_mesh on the r=20 sphere for several pipeline granularities (HCM_TRACE prints the host-side
timeline).  usage: python tools/e2e_pipeline.py [block,min_sub,threads ...]"""
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
ext = 2 * R + 3
configs = sys.argv[1:] or ["2048,512,15", "1024,512,15", "4096,512,15", "2048,512,8", "2048,512,4"]
os.environ["HCM_TRACE"] = "1"
for cfg in configs:
    blk, sub, thr = cfg.split(",")
    os.environ["HCM_UPLOAD_BLOCK"], os.environ["HCM_MIN_SUB_BATCH"], os.environ["HCM_HOST_THREADS"] = blk, sub, thr
    m = M.Mesher(origin=(-R - 1,) * 3, extent=(ext,) * 3, max_chunks=len(resident), max_tasks=len(meshed))
    ts = []
    for it in range(5):
        t0 = time.perf_counter(); r = m.upload_and_mesh(resident, pb.array, ph.array, meshed); ts.append(1e3 * (time.perf_counter() - t0))
    print(f"block {blk} min_sub {sub} threads {thr}: " + " ".join(f"{t:.1f}" for t in ts) + " ms", flush=True)
    # stage times alone, for comparison
    t0 = time.perf_counter(); m.upload(resident, pb.array, ph.array); m.sync(); t1 = time.perf_counter()
    m.mesh_async(meshed); m.wait(); t2 = time.perf_counter(); m.fetch(); t3 = time.perf_counter()
    print(f"   alone: upload {1e3*(t1-t0):.1f} mesh {1e3*(t2-t1):.1f} fetch {1e3*(t3-t2):.1f} ms", flush=True)
    m.close()
