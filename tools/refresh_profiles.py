"""Turn one capture set under gpurun_out/ (prefix given on the command line, e.g. r17) into the tracked summaries under
profiles/r01_end_* and profiles/ncu_summary.json.  Needs: <p>_launches_r20.csv, <p>_prof_mesh.ncu-rep, <p>_phase_timing.txt,
<p>_bench_full.json, and the libhcm.so the capture was taken with (for the SASS -> source line map)."""
import collections, csv, json, os, re, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pre = sys.argv[1]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
shutil.copy(f"{G}/{pre}_launches_r20.csv", f"{P}/r01_end_launches_r20_with_dram.csv")
shutil.copy(f"{G}/{pre}_bench_full.json", f"{P}/r01_end_bench_r20_n1.json")
shutil.copy(f"{G}/{pre}_phase_timing.txt", f"{P}/r01_end_mesh_kernel_phase_cycles.txt")
rep = f"{G}/{pre}_prof_mesh.ncu-rep"
det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout.split("\n")[:170]
open(f"{P}/r01_end_mesh_kernel_details.txt", "w").write("\n".join(det) + "\n")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
open("/tmp/_mesh_src.csv", "w").write(src)
os.makedirs("/tmp/_cub", exist_ok=True)
subprocess.run(["cuobjdump", "-xelf", "all", f"{ROOT}/hypercraft_b200/csrc/libhcm.so"], cwd="/tmp/_cub", capture_output=True)
dis = subprocess.run(["nvdisasm", "-g", "-c", "/tmp/_cub/hcm_kernels.sm_100a.cubin"], capture_output=True, text=True).stdout
open("/tmp/_k.txt", "w").write(dis)
# function / phase ranges from the current source
lines = open(f"{ROOT}/hypercraft_b200/csrc/hcm_kernels.cu").read().split("\n")
def find(pat, start=0):
    return next(i + 1 for i, l in enumerate(lines) if i >= start and re.search(pat, l))
marks = {int(m.group(1)): i + 1 for i, l in enumerate(lines) for m in [re.search(r"^\s+PHASE_MARK\((\d+)\);", l)] if m}
f = {n: find(p) for n, p in dict(texshow=r"bool tex_show\(", light4=r"void light4_corner\(", emit_help=r"void aabb_atomic\(", flow=r"DynRec flow_record\(",
                                  flood=r"u64 combine3\(", merge_warp=r"u32 pack_geo\(", merge_thread=r"void merge_mask_thread\(", emit=r"void emit_quads\(",
                                  custom=r"void emit_custom_voxel\(", kernel=r"^__global__ .*mesh_kernel\(").items()}
order = ["texshow", "light4", "emit_help", "flow", "flood", "merge_warp", "merge_thread", "emit", "custom", "kernel"]
rng = [("helpers", 1, 103)] + [(n, f[n], f[order[i + 1]] - 1) for i, n in enumerate(order[:-1])]
ph = [("p0_setup", f["kernel"], marks[1]), ("p4_light", marks[1] + 1, marks[17]), ("p1_planes", marks[17] + 1, marks[2]), ("p1b", marks[2] + 1, marks[3]),
      ("p2_faces", marks[3] + 1, marks[4]), ("p3", marks[4] + 1, marks[5]), ("p5", marks[5] + 1, marks[8]), ("p6a_fill", marks[8] + 1, marks[9]),
      ("p6b_keys", marks[9] + 1, marks[10]), ("p6c_merge_sched", marks[10] + 1, marks[11] - 3), ("p6c_merge_barrier", marks[11] - 2, marks[11]),
      ("p6d_scan_alloc", marks[11] + 1, marks[13]), ("p7_emit_custom_desc", marks[13] + 1, marks[16] + 40)]
spec = ",".join(f"{n}:{a}-{b}" for n, a, b in rng + ph)
out = subprocess.run([sys.executable, f"{ROOT}/tools/ncu_funcs.py", "/tmp/_mesh_src.csv", "/tmp/_k.txt", "mesh_kernel", spec], capture_output=True, text=True).stdout
open(f"{P}/r01_end_mesh_kernel_by_phase.txt", "w").write(out)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.split("\n")))
h, v = r[0], r[2]
want = {"gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__issue_active.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size"}
keep = [f"{a} {c}" for a, c in zip(h, v) if a in want or re.match(r"sm__inst_executed_pipe_(alu|fma|lsu|xu|adu)\.avg\.pct_of_peak_sustained_active$", a)
        or ("issue_stalled" in a and a.endswith("per_issue_active.ratio"))]
open(f"{P}/r01_end_mesh_kernel_counters.txt", "w").write("# ncu --set full, mesh_kernel, r=11 load sphere (1576 worked tasks), one launch\n" + "\n".join(sorted(keep)) + "\n")
rows = list(csv.reader(open(f"{G}/{pre}_launches_r20.csv")))
hd = next(i for i, x in enumerate(rows) if x and x[0] == "ID")
hh = rows[hd]
ki, mi, vi, ii = hh.index("Kernel Name"), hh.index("Metric Name"), hh.index("Metric Value"), hh.index("ID")
d = collections.OrderedDict()
for x in rows[hd + 1:]:
    if len(x) > vi:
        d.setdefault((int(x[ii]), x[ki].split("(")[0]), {})[x[mi]] = float(x[vi].replace(",", ""))
items = list(d.items())
i = [j for j, (k, m) in enumerate(items) if k[1] == "derive_kernel" and m["gpu__time_duration.sum"] > 5e5][-1]
der, cls, msh = items[i], items[i + 1], items[i + 2]
b = json.load(open(f"{P}/r01_end_bench_r20_n1.json"))
s = {"source": f"profiles/r01_end_launches_r20_with_dram.csv, launch ids {der[0][0]}-{msh[0][0]} = one timed step: derive {der[0][0]}, classify {cls[0][0]}, mesh {msh[0][0]} "
               "(ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv; python bench.py --steps 2 --warmup 1 --no-cpu-baseline; r=20)",
     "mesh_kernel_dram_bytes_per_launch": int(msh[1]["dram__bytes_read.sum"] + msh[1]["dram__bytes_write.sum"]),
     "mesh_kernel_duration_ns_under_ncu": msh[1]["gpu__time_duration.sum"],
     "derive_kernel_dram_bytes_per_launch": int(der[1]["dram__bytes_read.sum"] + der[1]["dram__bytes_write.sum"]),
     "derive_kernel_duration_ns_under_ncu": der[1]["gpu__time_duration.sum"],
     "classify_kernel_duration_ns_under_ncu": cls[1]["gpu__time_duration.sum"]}
tot = s["mesh_kernel_duration_ns_under_ncu"] + s["derive_kernel_duration_ns_under_ncu"] + s["classify_kernel_duration_ns_under_ncu"]
s["share_of_step"] = {k: round(s[f"{k}_kernel_duration_ns_under_ncu"] / tot, 4) for k in ("mesh", "derive", "classify")}
k = b["kernel_ms"]
t = k["mesh"] + k["derive"] + k["classify"]
s["note"] = ("per-launch times under ncu are cold-cache and serialised; compare shares, not absolutes. bench.py's live CUDA-event times for the same step "
             f"(profiles/r01_end_bench_r20_n1.json): mesh {k['mesh']:.2f} ms ({100 * k['mesh'] / t:.1f} %), derive {k['derive']:.2f} ms ({100 * k['derive'] / t:.1f} %), "
             f"classify {k['classify']:.3f} ms")
json.dump(s, open(f"{P}/ncu_summary.json", "w"), indent=1)
print(out.split("\n")[0])
print(s["share_of_step"], s["mesh_kernel_dram_bytes_per_launch"], s["derive_kernel_dram_bytes_per_launch"])
print({q: b[q] for q in ("value", "ms_per_step", "kernel_ms", "quads_per_s")}, b["e2e"]["ms_per_step"], b["roofline"]["frac"], b["roofline"]["achieved"], b["cpu_baseline"]["value"])
