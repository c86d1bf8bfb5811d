"""Developer tool: upload the r = $R load sphere (reference terrain) and run three derive + mesh steps -- the target of the ncu --set full
captures in tools/capture_profiles.sh.  HCM_SO selects another build of the library."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import hcoracle
from hypercraft_b200 import mesher as M, worldgen
R = int(os.environ.get("R", "11"))
meshed, resident = worldgen.load_radius_positions(R)
ref = hcoracle.RefWorld(); ref.generate(resident)
b, h = ref.get_chunks(resident)
ext = 2 * R + 3
if os.environ.get("HCM_SO"): M.LIB_PATH = os.environ["HCM_SO"]
m = M.Mesher(origin=(-R - 1,) * 3, extent=(ext,) * 3, max_chunks=len(resident), max_tasks=len(meshed))
m.upload(resident, b, h); m.sync()
for _ in range(3):  # three steps (derive, classify, mesh): `ncu -k regex:<kernel> -s 1 -c 1` takes the second launch of either kernel
    m.rederive(); m.sync()
    m.mesh_async(meshed); m.wait()
st = m.stats()
print("chunks", len(resident), {k: st[k] for k in st if k.endswith("_ms")})
