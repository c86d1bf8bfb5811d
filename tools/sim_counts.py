"""Developer tool: algorithmic event counts of mesh_kernel on reference terrain, taken on the CPU emulation build with
-DHCM_PHASE_TIMING (tests/cusim; `make -C tests/cusim OUT=_build_count SIMFLAGS=-DHCM_PHASE_TIMING`).  Counts, not times: the
emulation has no clock.  R=<radius> (default 6)."""
import ctypes, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
sim = os.path.join(ROOT, "tests", "cusim")
subprocess.run(["make", "-C", sim, "OUT=_build_count", "SIMFLAGS=-DHCM_PHASE_TIMING"], check=True, stdout=subprocess.DEVNULL)
import hcoracle
from hypercraft_b200 import mesher as M, worldgen
M.LIB_PATH = os.path.join(sim, "_build_count", "libhcm_sim.so")
R = int(os.environ.get("R", "6"))
meshed, resident = worldgen.load_radius_positions(R)
ref = hcoracle.RefWorld(); ref.generate(resident)
b, h = ref.get_chunks(resident)
ext = 2 * R + 3
m = M.Mesher(origin=(-R - 1,) * 3, extent=(ext,) * 3, max_chunks=len(resident), max_tasks=len(meshed))
m.set_option("tune", 0)  # no A/B rounds here: one form per run
if os.environ.get("V"):  # V=1: the variant forms (include/hcm.h, "Kernel variants")
    m.set_option("derive_variant", int(os.environ["V"])), m.set_option("mesh_variant", int(os.environ["V"]))
m.upload(resident, b, h)
L = M.load_library()
out = (ctypes.c_ulonglong * 64)()
L.hcm_debug_phase_cycles(out, 1)
m.mesh_async(meshed); c = m.wait()
L.hcm_debug_phase_cycles(out, 0)
n = c["n_worked"]
print(f"r={R}: {len(meshed)} tasks, {n} worked, {c['n_vertices'] // 4} quads")
print(f"  flood: {out[28]/n:.2f} steps/task, {out[27]/n:.0f} column evaluations/task ({out[27]/max(out[28],1):.0f} per step), {out[25]/n:.0f} of them gained cells ({100*out[25]/max(out[27],1):.1f} %)")
print(f"  warp merges: {out[20]/n:.1f} masks/task, {out[22]/n:.0f} rows/task, {out[39]/n:.0f} row iterations/task, {out[55]/max(out[39],1):.2f} height steps/iteration in {out[57]/max(out[39],1):.2f} dependent round trips; tiny-mask warps {out[36]/n:.2f}/task; quads {out[23]/n:.0f}/task")
for i in range(64):
    if out[i] and i not in (20, 22, 23, 25, 27, 28, 36, 39, 55, 57):
        print(f"  counter {i}: {out[i]/n:.1f}/task")
