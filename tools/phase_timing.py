"""Developer tool: per-phase cycle breakdown of mesh_kernel (builds csrc/libhcm_timing.so with -DHCM_PHASE_TIMING)."""
import ctypes, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
csrc = os.path.join(ROOT, "hypercraft_b200", "csrc")
so = os.environ.get("HCM_TIMING_SO") or os.path.join(csrc, "libhcm_timing.so")
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
out = (ctypes.c_ulonglong * 64)()
L.hcm_debug_phase_cycles(out, 1)
m.mesh_async(meshed); c = m.wait(); st = m.stats()
L.hcm_debug_phase_cycles(out, 0)
names = ["task fetch", "task set-up", "planes", "liquid", "faces", "counts+prefix", "flood setup", "flood+unpack", "custom count", "fill", "keys", "merge",
         "scan", "alloc", "emit", "custom emit", "descriptors", "TMA wait"]
tot = sum(out[:18]); n = c["n_worked"]
print(f"r={R}: {n} worked tasks, mesh_ms {st['mesh_ms']:.3f}, cycles/task {tot / n:.0f}")
print(f"  per task: masks merged {out[20]/n:.1f}, faces {out[21]/n:.0f}, non-empty rows {out[22]/n:.0f}, quads {out[23]/n:.0f}")
print("  masks per task by face count (<=8,<=16,<=32,<=64,<=128,>128):", [round(out[40+i]/n,1) for i in range(6)], " rows:", [round(out[48+i]/n,1) for i in range(6)])
print(f"  flood: steps executed per task {out[28]/n:.1f} (sun+torch), live column evaluations per task {out[27]/n:.0f} (of {15*3424} possible per channel)")
print(f"  flood: step 0 {out[31]/n:.0f} cyc/task, later steps {out[32]/max(out[28]-2*n,1):.0f} cyc/step")
ls_ = max(out[28]-n,1)
print(f"  flood later steps (tid 0): dilation {out[52]/ls_:.0f}, inject+planes+publish {out[53]/ls_:.0f}, barrier wait {out[54]/ls_:.0f} cycles/step ({ls_/n:.1f} later steps/task)")
print(f"  flood step loop {out[29]/n:.0f} cyc/task, unpack {out[30]/n:.0f} cyc/task")
print(f"  exact Show() candidate pairs per task: {out[26]/n:.0f}")
print(f"  warp-level merges: {out[20]/n:.1f} masks/task, {out[22]/n:.0f} rows/task, {out[24]/max(out[22],1):.0f} cycles/row, longest single mask {out[33]} cycles")
print(f"  warp merge: set-up {out[37]/max(out[20],1):.0f} cyc/mask, main loop {out[38]/max(out[39],1):.0f} cyc/iteration, {out[39]/n:.0f} iterations/task")
print(f"  warp merge: height steps {out[55]/max(out[39],1):.1f} per iteration, record pass {out[56]/max(out[20],1):.0f} cyc/mask")
print(f"  flood set-up (thread 64): wait for task set-up {out[61]/n:.0f}, zeroing {out[62]/n:.0f}, loads {out[63]/n:.0f}, barrier after loads {out[60]/n:.0f} cycles/task")
print(f"  tiny-mask warps: {out[36]/n:.2f} per task, {out[35]/max(out[36],1):.0f} cycles each; merge section per-warp mean {out[34]/n/32:.0f} cycles")
for i, nm in enumerate(names):
    print(f"  {nm:14s} {out[i] / n:9.0f} cyc/task  {100 * out[i] / tot:5.1f}%")
