"""Developer tool: per-phase cycle breakdown of mesh_kernel (builds csrc/libhcm_timing.so with -DHCM_PHASE_TIMING)."""
import ctypes, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
csrc = os.path.join(ROOT, "hypercraft_b200", "csrc")
so = os.path.join(csrc, "libhcm_timing.so")
if "--build" in sys.argv or not os.path.exists(so):
    subprocess.run(["nvcc", "-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC,-fvisibility=hidden",
                    "-DHCM_PHASE_TIMING", "-shared", "-o", so, os.path.join(csrc, "hcm_api.cu"), os.path.join(csrc, "hcm_kernels.cu")], check=True)
    if "--build" in sys.argv: sys.exit(0)
import hcoracle
from hypercraft_b200 import mesher as M, worldgen
M.LIB_PATH = so
R = int(os.environ.get("R", "11"))
meshed, resident = worldgen.load_radius_positions(R)
ref = hcoracle.RefWorld(); ref.generate(resident)
b, h = ref.get_chunks(resident)
ext = 2 * R + 3
m = M.Mesher(origin=(-R - 1,) * 3, extent=(ext,) * 3, max_chunks=len(resident), max_tasks=len(meshed))
m.upload(resident, b, h)
for _ in range(3):
    m.mesh_async(meshed); c = m.wait()
L = M.load_library()
out = (ctypes.c_ulonglong * 32)()
L.hcm_debug_phase_cycles(out, 1)
m.mesh_async(meshed); c = m.wait(); st = m.stats()
L.hcm_debug_phase_cycles(out, 0)
names = ["task fetch", "TMA+zero", "planes", "liquid", "faces", "counts+prefix", "flood setup", "flood+unpack", "custom count", "fill", "keys", "merge",
         "scan", "alloc", "emit", "custom emit", "descriptors"]
tot = sum(out[:17]); n = c["n_worked"]
print(f"r={R}: {n} worked tasks, mesh_ms {st['mesh_ms']:.3f}, cycles/task {tot / n:.0f}")
print(f"  per task: masks merged {out[20]/n:.1f}, faces {out[21]/n:.0f}, non-empty rows {out[22]/n:.0f}, quads {out[23]/n:.0f}")
print(f"  exact Show() candidate pairs per task: {out[26]/n:.0f}")
print(f"  merge_mask cycles summed over warps {out[24]/n:.0f}/task ; per-warp busy-until sum {out[25]/n:.0f}/task (16 warps)")
for i, nm in enumerate(names):
    print(f"  {nm:14s} {out[i] / n:9.0f} cyc/task  {100 * out[i] / tot:5.1f}%")
