"""Developer tool: what can be said about the kernels WITHOUT a GPU -- ptxas resource usage (registers, spills, static shared memory) and
SASS instruction counts per kernel of hypercraft_b200/csrc/libhcm.so, with the mnemonics that show TMA / mbarrier / warp-collective use
(B200_PROFILING.md names them).  Static facts, not measurements.   python tools/static_kernel_table.py > profiles/<round>_static_kernel_table.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "hypercraft_b200", "csrc")
WATCH = ["UTMALDG", "UBLKCP", "UBLKPF", "SYNCS", "BAR", "REDUX", "SHFL", "VOTE", "MATCH", "ATOMS", "ATOMG", "RED", "LDS", "STS", "LDG", "STG", "LDL", "STL", "POPC", "FLO"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    subprocess.run(["make", "-C", CSRC], check=True, stdout=subprocess.DEVNULL)
    log = open(os.path.join(CSRC, "ptxas.log")).read()
    res = {}
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                         r"ptxas info\s+: Used (\d+) registers, used (\d+) barriers(?:, \d+ bytes cumulative stack size)?(?:, (\d+) bytes smem)?", log):
        res[m.group(1)] = dict(stack=int(m.group(2)), spill_st=int(m.group(3)), spill_ld=int(m.group(4)), regs=int(m.group(5)), smem=int(m.group(7) or 0))
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(CSRC, "libhcm.so")], capture_output=True, text=True).stdout
    counts = {}
    for part in re.split(r"\n\s*Function : ", sass)[1:]:
        name = part.split("\n", 1)[0].strip()
        c = collections.Counter()
        n = 0
        for line in part.split("\n"):
            mm = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if mm:
                n += 1
                op = mm.group(1)
                for w in WATCH:
                    if op == w or op.startswith(w + "."):
                        c[w] += 1
        counts[name] = (n, c)
    names = demangle(sorted(counts))
    print("Static facts about the kernels of libhcm.so (nvcc 12.9, -O3, sm_100a): ptxas resources and SASS instruction counts.")
    print("NOT measurements -- round 2 had no GPU.  UTMALDG = cp.async.bulk.tensor (3-D TMA box), UBLKCP = cp.async.bulk (1-D TMA),")
    print("UBLKPF = cp.async.bulk.prefetch.L2, SYNCS = mbarrier operations, REDUX / SHFL / VOTE / MATCH = warp collectives.\n")
    for k in sorted(counts, key=lambda k: names[k]):
        n, c = counts[k]
        r = res.get(k, {})
        short = re.sub(r"\(.*", "", names[k]).replace("hcm::", "").replace("void ", "")
        print(f"{short}")
        print(f"    {n} SASS instructions, {r.get('regs', '?')} registers, {r.get('spill_st', '?')} B spill stores / {r.get('spill_ld', '?')} B spill loads, "
              f"{r.get('smem', 0)} B static smem")
        print("    " + "  ".join(f"{w} {c[w]}" for w in WATCH if c[w]))
    print("\nderive_kernel<0, false> and mesh_kernel<0> are the round-1 kernels (same instruction streams as the r01_head captures describe;")
    print("mesh_kernel<0> differs by the position of one IMAD.MOV).  The <k, true> forms are the A/B rounds' trial launches (bounded waits).")


if __name__ == "__main__":
    main()
