#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 chunk mesher (metric and config from BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W]                    # this repo's CUDA path
  python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]   # the reference's own CPU mesher (oracle/_ref)

Workload (BASELINE.json configs[1]): full load-radius remesh -- every chunk p with |p|^2 <= r^2 around a centre is
meshed, its 26 neighbours coming from the one-chunk dilation of that sphere (misc/load_radius.cpp:16-35), r = 20 =
kWorldMaxLoadRadius (client/include/client/Config.hpp:14), terrain from the reference's DefaultTerrain(12314524)
(client/src/LocalClient.cpp:37).  One STEP = one full remesh of that world:
    hcm_rederive_chunks(all resident) + hcm_mesh_batch(all tasks)            -> `value`   (inputs resident in HBM)
    hcm_upload_and_mesh(host chunks in, host meshes out)                     -> `e2e`     (host buffers both sides;
                                                  = hcm_upload_chunks + hcm_mesh_batch + hcm_mesh_fetch, overlapped)
Multi-GPU (--gpus N under torchrun): chunks are independent given their halo, so every rank meshes its own sphere
(centre shifted by 3r chunks per rank) with no collective on the data path -> weak scaling; value = total chunks of
all ranks / max-over-ranks time.  For N > 1 the line also carries `config4`: ONE r = 32 sphere (BASELINE.json
configs[3]) cut into x-slabs balanced by estimated worked chunks, one slab (dilated by a chunk) per GPU, kernel-only and
end to end.  For N = 1 it carries `secondary_configs`: configs[2] (worst-case chunks) and configs[4] (edit-driven remesh).

Parity gate: EVERY chunk meshed in the last end-to-end step of EVERY rank is compared with the reference CPU mesher through
a digest of its serialised meshes (mesh count, flags, AABBs, vertex bytes, index bytes); the mismatch count is summed over
the ranks.  `parity.byte_compared_chunks == parity.chunks_meshed` is the bar.

`kernel_variants`: derive_kernel and mesh_kernel ship in a baseline form and in re-scheduled variant forms; a context A/Bs them on its
device during its first large calls (here: the warm-up steps) and runs what that selected in the timed steps (include/hcm.h, "Kernel
variants").  The block reports the verdicts, the mismatch counts and every form's CUDA-event time.

A watchdog guards every stage: if an optional block (device sink, hinted upload, round-1 upload mode, config 4, secondary configs, the
sharded-dispatcher check) hangs or fails, the line is still printed from what is complete, with the stage named under "watchdog" / "error".

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

SEED = 12314524  # client/src/LocalClient.cpp:37
H_BYTES = 34 ** 3 * 3  # blocks (2 B) + light (1 B) on [-1,32]^3: SURVEY.md section 8d
# derive_kernel per chunk: reads Block data + heights, writes property bytes, pass / torch words, heights
DERIVE_READ, DERIVE_WRITE_FIXED, DERIVE_WRITE_PER_LEVEL = 65536 + 1024, 32768 + 4096 + 1024, 4096


class Watchdog:
    """A bench line must come out even if an optional block (device sink, hinted upload, config 4, secondary configs) hangs or a rank
    dies inside a collective: every stage arms a deadline; when one passes, rank 0 prints the line from what is complete (`line_fn`,
    set once the headline numbers exist) with the stage named under "watchdog", and every rank leaves.  Without a headline the exit
    code is 3 -- a failed run, not a hung box."""

    def __init__(self, rank: int):
        self.rank, self.stage, self.deadline, self.line_fn, self.lock, self.printed = rank, "start", None, None, threading.Lock(), False
        threading.Thread(target=self._run, daemon=True).start()

    def arm(self, stage: str, seconds: float):
        self.stage, self.deadline = stage, time.time() + seconds

    def emit(self, line: dict):
        with self.lock:
            if not self.printed:
                self.printed = True
                print(json.dumps(line), flush=True)

    def _run(self):
        while True:
            time.sleep(1.0)
            if self.deadline is not None and time.time() > self.deadline:
                msg = f"stage '{self.stage}' exceeded its time budget; the line holds what was complete"
                sys.stderr.write(f"bench.py[rank {self.rank}]: watchdog: {msg}\n")
                sys.stderr.flush()
                if self.line_fn is not None:
                    if self.rank == 0:
                        try:
                            self.emit({**self.line_fn(), "watchdog": msg})
                        except Exception as e:  # never hang in the handler
                            sys.stderr.write(f"bench.py: watchdog could not build the line: {e}\n")
                            os._exit(3)
                    os._exit(0)
                os._exit(3)


def guarded(what: str, fn, *a, **kw):
    """Optional blocks never take the headline down: an exception becomes {"error": ...}."""
    try:
        return fn(*a, **kw)
    except BaseException as e:  # noqa: BLE001 -- incl. SystemExit from a failed native call wrapper
        if isinstance(e, KeyboardInterrupt):
            raise
        return {"error": f"{what}: {type(e).__name__}: {e}"[:600]}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--radius", type=int, default=20)
    ap.add_argument("--strong", action="store_true",
                    help="make BASELINE config 4 the headline: ONE load sphere cut into x-slabs, one slab (dilated by a chunk) per GPU")
    ap.add_argument("--config4-radius", type=int, default=32, help="radius of the config4 block printed for --gpus N > 1 (0 = skip it)")
    ap.add_argument("--no-secondary", action="store_true", help="skip secondary_configs (configs 3 and 5) at N = 1")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU work per step of the --impl reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU leg (parity gate and cpu_baseline): profiling runs only")
    ap.add_argument("--dry-run-emulated", action="store_true",
                    help="DEVELOPER DRY RUN, never a measurement: run this script's own logic against tests/cusim (the CPU emulation build of the "
                         "library, test infrastructure) on a machine without a GPU; the line says \"emulated\": true and its numbers mean nothing")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def make_cpu_arm(threads: int):
    """(arm, provide(positions), data label): the CPU side of the bench -- input generator, parity checker and CPU baseline.  The
    unmodified reference from oracle/_ref when it was built, else this repo's port on its own integer-hash terrain."""
    import hcoracle
    from hypercraft_b200 import worldgen

    if use_reference_build():
        ref = hcoracle.RefWorld()

        def provide(need):
            ref.generate(need, seed=SEED, threads=threads)

        return ref, provide, f"synthetic: reference DefaultTerrain(seed {SEED}) run on this host via oracle/_ref (no dataset)"
    arm = PortArm({})

    def provide(need):
        for q in need:
            key = tuple(int(c) for c in q)
            if key not in arm.chunks:
                arm.chunks[key] = worldgen.synth_terrain_chunk(SEED, q)

    return arm, provide, "synthetic: repo height-field terrain (oracle/_ref terrain generator not available)"


def fill_pinned(arm, resident):
    """Pinned host arrays (blocks, heights) of `resident` from the CPU arm's world."""
    from hypercraft_b200 import mesher

    pb = mesher.PinnedArray((len(resident), 32768), np.uint16)
    ph = mesher.PinnedArray((len(resident), 1024), np.int8)
    if isinstance(arm, PortArm):
        for i, p in enumerate(resident):
            pb.array[i], ph.array[i] = arm.chunks[tuple(int(c) for c in p)]
    else:
        arm.get_chunks(resident, pb.array, ph.array)
    return pb, ph


def workload_name(radius: int, strong: bool = False) -> str:
    if strong:
        return f"full load-radius remesh r={radius} (misc/load_radius.cpp), ONE sphere cut into x-slabs across the GPUs (BASELINE config 4)"
    return f"full load-radius remesh r={radius} (misc/load_radius.cpp; kWorldMaxLoadRadius) per GPU"


def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def algorithmic_bytes(n_chunks: int, n_vertices: int, n_meshes: int) -> dict:
    """SURVEY 8d for `n_chunks` chunks that emit `n_vertices` vertices: the canonical figure roofline.achieved uses (34^3 halo blocks
    2 B + light 1 B per chunk, 8 B per vertex written) and the two secondary ones -- "full": what the reference task really touches
    (62^3 blocks + 3 x 62^2 heights, light re-derived from raw blocks); "min": every chunk read once from HBM with perfect halo reuse,
    vertices + indices + mesh infos."""
    return {"canonical": n_chunks * H_BYTES + 8 * n_vertices,
            "full": n_chunks * (62 ** 3 * 2 + 62 * 62 * 3) + 8 * n_vertices,
            "min": n_chunks * 66560 + 11 * n_vertices + 40 * n_meshes}


def roofline_secondary(algo: dict, kernel_ms: float, peak_gbs: float) -> dict:
    """The two secondary byte models of SURVEY 8d against the same launch time and peak as roofline.achieved."""
    return {"secondary": {k: {"algorithmic_bytes_per_launch": int(algo[k]), "achieved": algo[k] / (kernel_ms / 1e3) / 1e9,
                              "frac": algo[k] / (kernel_ms / 1e3) / 1e9 / peak_gbs} for k in ("full", "min")},
            "secondary_formulas": {"full": "62^3*2 + 62*62*3 + 8*V per chunk (what the reference task touches)",
                                   "min": "66560 + 11*V + 40*M per chunk (each chunk read once, vertices + indices + mesh infos)"}}


def derive_bytes(n_chunks: int, n_levels: int = 1) -> int:
    """derive_kernel's algorithmic bytes for one launch over n_chunks resident chunks."""
    return n_chunks * (DERIVE_READ + DERIVE_WRITE_FIXED + DERIVE_WRITE_PER_LEVEL * n_levels)


def roofline_block(n_meshed: int, n_worked: int, n_resident: int, n_vertices: int, n_meshes: int, kernel_ms: dict, ms_per_step: float, peak_gbs: float,
                   peak_src: str, traffic: dict | None) -> dict:
    """The `roofline` object of the bench line.  The dominant kernel is mesh_kernel, and ONE launch of it processes the `n_worked`
    chunks classify_kernel did not answer from the chunk flags -- those are its units (every emitted vertex belongs to one of
    them).  `step_frac` is the whole step (derive + classify + mesh) against the canonical bytes of ALL meshed chunks."""
    k_ms = kernel_ms["mesh"]
    algo = algorithmic_bytes(n_worked, n_vertices, n_meshes)
    achieved = algo["canonical"] / (k_ms / 1e3) / 1e9
    all_bytes = algorithmic_bytes(n_meshed, n_vertices, n_meshes)["canonical"]
    d_bytes = derive_bytes(n_resident)
    d_ach = d_bytes / (kernel_ms["derive"] / 1e3) / 1e9
    t = traffic or {}
    return {"bound": "hbm", "kernel": "hcm::mesh_kernel", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
            "traffic": t.get("mesh_kernel_dram_bytes_per_launch"), "traffic_source": t.get("source"), "peak_source": peak_src,
            "units_per_launch": int(n_worked), "unit_name": "chunks through the full light + mesh path (classify_kernel answers the others from chunk flags)",
            "algorithmic_bytes_per_launch": int(algo["canonical"]), "formula": "sum over the launch's worked chunks of 34^3*3 + 8*vertices (SURVEY 8d)",
            "kernel_ms": k_ms,
            "step_frac": all_bytes / (ms_per_step / 1e3) / 1e9 / peak_gbs,
            "step_formula": "sum over ALL meshed chunks of 34^3*3 + 8*vertices, over ms_per_step (derive + classify + mesh)",
            "derive_kernel": {"kernel": "hcm::derive_kernel", "units_per_launch": int(n_resident), "algorithmic_bytes_per_launch": int(d_bytes),
                              "formula": "per resident chunk: 66560 read + 41984 written (property bytes, pass + torch words, heights)",
                              "kernel_ms": kernel_ms["derive"], "achieved": d_ach, "frac": d_ach / peak_gbs,
                              "traffic": t.get("derive_kernel_dram_bytes_per_launch")},
            **roofline_secondary(algo, k_ms, peak_gbs)}


def pool_baseline(ref, positions, cpu_threads: int) -> dict:
    """The CPU arm at the reference client's own worker count (Application.cpp:248: hardware_concurrency() * 3 / 4), same sample."""
    pool_threads = max(1, cpu_threads * 3 // 4)
    st, _ = ref.mesh_batch(positions, threads=pool_threads)
    return {"cores": pool_threads, "value": st["chunks"] / st["seconds"], "unit": "chunks/s",
            "note": "hardware_concurrency()*3/4 workers, the reference client's choice (Application.cpp:248)"}


class PortArm:
    """The CPU arm when oracle/_ref is absent: this repo's restatement (oracle/hc_oracle.cpp, kind "port") behind the two
    calls bench.py makes on a RefWorld -- mesh_batch and mesh -- over chunks held in a dict."""
    kind = "port"
    what = "oracle/hc_oracle.cpp (CPU restatement of ChunkTaskRunner<kMesh>::Run), one call stream per thread"

    def __init__(self, chunks: dict):
        import hcoracle
        from hypercraft_b200 import registry

        self.port = hcoracle.PortOracle(registry.load().blob)
        self.chunks = chunks

    def mesh_batch(self, positions, threads=1, per_chunk=False):
        st, per, _ = self.port.mesh_batch(self.chunks, positions, threads=threads)
        return st, (per if per_chunk else None)

    def mesh(self, pos):
        return self.port.mesh(self.chunks, tuple(int(c) for c in pos))


def parse_cpulist(text: str) -> list:
    """"0-15,32-47" -> [0..15, 32..47] (the format of sysfs cpulist files)."""
    out = []
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out.extend(range(int(a), int(b or a) + 1))
    return out


def split_cores_by_gpu(allowed: list, local_lists: list) -> list:
    """Cores per rank for ranks sharing one host.  `local_lists[r]` = the cores on the NUMA node of rank r's GPU (sysfs
    local_cpulist; None if unknown), `allowed` = this process' affinity mask.  Ranks whose GPUs hang off the same node share that node's
    cores in rank order, so that a rank's scan threads, its pinned buffers (first touch) and its GPU's PCIe root sit on one socket; a rank
    whose node is unknown, or whose share would be empty, falls back to an even split of the whole mask."""
    world = len(local_lists)
    allowed = sorted(allowed)
    per = max(1, len(allowed) // world)
    even = [allowed[r * per:(r + 1) * per] or allowed for r in range(world)]
    out = []
    for r, mine in enumerate(local_lists):
        share = None
        if mine:
            node = [c for c in sorted(mine) if c in set(allowed)]
            peers = [q for q in range(world) if local_lists[q] and sorted(local_lists[q]) == sorted(mine)]
            k, n = peers.index(r), len(peers)
            width = len(node) // n
            share = node[k * width:(k + 1) * width] if width else None
        out.append(share or even[r])
    return out


def pin_rank_to_cores(local: int, world: int, emulated: bool) -> dict:
    """sched_setaffinity of this rank to its share of the host's cores (see split_cores_by_gpu); -> what was done, for the line."""
    try:
        allowed = sorted(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        return {"cores": None, "how": "sched_getaffinity unavailable"}
    lists = [None] * world
    how = "even split of the affinity mask"
    if not emulated:
        try:
            import torch

            for r in range(world):
                p = torch.cuda.get_device_properties(r)
                path = f"/sys/bus/pci/devices/{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0/local_cpulist"
                lists[r] = parse_cpulist(open(path).read()) or None
            if all(lists):
                how = "cores of the NUMA node of each rank's GPU (sysfs local_cpulist), shared among the ranks of that node"
        except Exception:  # noqa: BLE001 -- no sysfs entry, fewer visible devices than ranks, older torch: even split
            lists = [None] * world
    mine = split_cores_by_gpu(allowed, lists)[local]
    try:
        os.sched_setaffinity(0, mine)
    except OSError:
        return {"cores": None, "how": "sched_setaffinity refused"}
    return {"cores": mine, "how": how}


def use_reference_build() -> bool:
    """oracle/_ref when it was compiled (it travels to the GPU box); HCM_BENCH_FORCE_PORT=1 selects the port (test knob)."""
    import hcoracle

    return hcoracle.have_ref() and not os.environ.get("HCM_BENCH_FORCE_PORT")


def compare_all_chunks(res, arm, meshed, threads: int) -> tuple[dict, dict]:
    """The parity gate: every task of the fetched batch `res` against the CPU arm, digest for digest (hcoracle.digest_batch serialises
    each task's meshes exactly like oracle/ref_shim.cpp does for the reference and hashes the bytes).  -> (parity dict, CPU stats)."""
    import hcoracle

    st, per = arm.mesh_batch(meshed, threads=threads, per_chunk=True)
    got = hcoracle.digest_batch(res, threads)
    count_bad = int((got[:, :3] != per[:, :3]).any(axis=1).sum())
    byte_bad = int((got[:, :4] != per[:, :4]).any(axis=1).sum())
    return ({"chunks_meshed": int(len(meshed)), "byte_compared_chunks": int(len(meshed)), "count_mismatches": count_bad, "byte_mismatches": byte_bad,
             "how": "FNV-1a of each chunk's serialised meshes (mesh count, flags, AABBs, vertex bytes, index bytes) vs the CPU arm's digest of the same chunk"},
            st)


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path (unmodified ChunkTaskRunner<kMesh>::Run from oracle/_ref, all host
    threads); the oracle port (cpu_baseline.kind "port") where oracle/_ref was not built."""
    if rank != 0:
        return
    from hypercraft_b200 import worldgen

    threads = os.cpu_count() or 1
    radius = args.radius
    meshed, resident = worldgen.load_radius_positions(radius)
    ref, provide, data = make_cpu_arm(threads)
    kind = getattr(ref, "kind", "reference")
    # bounded sample: a stride of the task list
    probe = meshed[:: max(1, len(meshed) // 256)][:256]
    provide(np.unique(np.concatenate([worldgen.neighbourhood(p) for p in probe]), axis=0))
    st, _ = ref.mesh_batch(probe, threads=threads)
    rate = st["chunks"] / max(st["seconds"], 1e-6)
    per_step = int(min(len(meshed), max(256, rate * args.cpu_seconds / max(1, args.steps + args.warmup) * 2)))
    sample = meshed[:: max(1, len(meshed) // per_step)][:per_step]
    provide(np.unique(np.concatenate([worldgen.neighbourhood(p) for p in sample]), axis=0))
    for _ in range(args.warmup):
        ref.mesh_batch(sample, threads=threads)
    secs, quads = 0.0, 0
    for _ in range(args.steps):
        st, _ = ref.mesh_batch(sample, threads=threads)
        secs += st["seconds"]
        quads = st["vertices"] // 4
    ms = secs / args.steps * 1e3
    value = len(sample) / (ms / 1e3)
    sample_txt = f"{len(sample)} of the {len(meshed)} chunks of the r={radius} load sphere (every {max(1, len(meshed) // per_step)}th), per step"
    out = {"impl": "reference", "metric": "chunks meshed/sec", "value": value, "unit": "chunks/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u16 integer",
           "data": data,
           "config": {"workload": workload_name(radius), "radius": radius, "chunks_per_step": len(sample),
                      "note": "each step meshes a bounded sample of the workload's task list (cpu_baseline.sample)"},
           "quads_per_s": quads / (ms / 1e3),
           "cpu_baseline": {"value": value, "unit": "chunks/s", "cores": threads, "kind": kind, "sample": sample_txt, "cpu_model": cpu_model()},
           "e2e": {"value": value, "unit": "chunks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


class Dist:
    """torch.distributed plumbing of the bench: barrier and small reductions only -- nothing of the data path.
    `emulated` (--dry-run-emulated): gloo and CPU tensors, so that the N > 1 logic can be exercised without GPUs."""

    def __init__(self, world: int, local: int, emulated: bool = False):
        import torch

        self.torch, self.world, self.emulated = torch, world, emulated
        self.device = "cpu" if emulated else "cuda"
        if world > 1:
            import torch.distributed as dist

            self.dist = dist
            if emulated:
                dist.init_process_group("gloo")
            else:
                dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def sync(self):
        if not self.emulated:
            self.torch.cuda.synchronize()

    def barrier(self):
        self.sync()
        if self.world > 1:
            self.dist.barrier()
        self.sync()

    def reduce(self, values, op: str = "sum") -> list:
        t = self.torch.tensor([float(v) for v in values], dtype=self.torch.float64, device=self.device)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return [float(x) for x in t.tolist()]

    def gather(self, value: float) -> list:
        t = self.torch.zeros(self.world, dtype=self.torch.float64, device=self.device)
        t[int(os.environ.get("RANK", "0"))] = float(value)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(x) for x in t.tolist()]

    def stream_timer(self, stream_ptr: int, local: int):
        """(start, stop -> ms): CUDA events recorded on the ctx's own stream (torch.cuda.Event only sees streams torch knows, so the
        ctx stream is wrapped as an ExternalStream); host clock in the emulated dry run."""
        if self.emulated:
            box = {}
            return (lambda: box.__setitem__("t", time.perf_counter())), (lambda: (time.perf_counter() - box["t"]) * 1e3)
        torch = self.torch
        stream = torch.cuda.ExternalStream(stream_ptr, device=torch.device("cuda", local))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def stop():
            e1.record(stream)
            self.sync()
            return e0.elapsed_time(e1)

        return (lambda: e0.record(stream)), stop

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def measure(D: Dist, local: int, meshed, resident, pb, ph, origin, extent, steps: int, warmup: int, arm, cpu_threads: int, check: bool,
            on_core=None) -> dict:
    """One workload on this rank's GPU: `steps` device-timed steps with the inputs resident in HBM, `steps` end-to-end steps from
    pinned host buffers, then the parity gate over every chunk of the last end-to-end result.  Times are max over ranks, counts are
    sums over ranks.  `on_core(out)` is called once those numbers exist, before the optional end-to-end variants run."""
    from hypercraft_b200 import mesher as M

    m = M.Mesher(origin=origin, extent=extent, device=local, max_chunks=len(resident), max_tasks=len(meshed))
    m.upload(resident, pb.array, ph.array)
    m.sync()
    t_start, t_stop = D.stream_timer(m.stream_ptr(), local)

    def step_resident():
        m.rederive()
        m.mesh_async(meshed)
        return m.wait()

    counts = None
    mesh_ms, derive_ms, classify_ms = [], [], []
    for _ in range(max(warmup, 3)):  # warm-up (also sizes the output arenas); the per-kernel event times are read here, one step at a time
        counts = step_resident()
        s = m.stats()
        mesh_ms.append(s["mesh_ms"]), derive_ms.append(s["derive_ms"]), classify_ms.append(s["classify_ms"])
    mesh_ms, derive_ms, classify_ms = mesh_ms[1:], derive_ms[1:], classify_ms[1:]  # the first step grows the arenas
    # ---- timed: K steps, device timed on the ctx stream, max over ranks --------------------------------------------
    # The K steps are enqueued back to back -- hcm_rederive_chunks and hcm_mesh_batch only enqueue -- and the host waits once, after the
    # last: the stream never idles for a host round trip between steps (the arenas were sized by the warm-up, every step does the same work).
    D.barrier()
    launches0 = m.stats()["kernel_launches"]
    t_start()
    for _ in range(steps):
        m.rederive()
        m.mesh_async(meshed)
    elapsed_ms = t_stop()
    counts = m.wait()
    D.barrier()
    s = m.stats()  # the ctx's own events around the kernels of the LAST timed step (the earlier steps' recordings were overwritten)
    warm = {"derive": float(np.mean(derive_ms)), "classify": float(np.mean(classify_ms)), "mesh": float(np.mean(mesh_ms))}
    mesh_ms, derive_ms, classify_ms = [s["mesh_ms"]], [s["derive_ms"]], [s["classify_ms"]]
    launches = s["kernel_launches"] - launches0
    variants = m.tune_info()  # the A/B rounds ran in the warm-up steps; the timed steps used what they selected
    ms_per_step = D.reduce([elapsed_ms], "max")[0] / steps
    tot_chunks, tot_quads, tot_launches, tot_worked = D.reduce([len(meshed), counts["n_vertices"] // 4, launches, counts["n_worked"]])

    # ---- end to end: host buffers in, host meshes out, every step ---------------------------------------------------
    D.barrier()
    for it in range(3):  # the three calls one after the other (first pass: warm-up of the pinned result buffers)
        if it == 1:
            t0 = time.perf_counter()
        m.upload(resident, pb.array, ph.array)
        res = m.mesh(meshed)
    e2e_serial_ms = (time.perf_counter() - t0) * 1e3 / 2
    res = m.upload_and_mesh(resident, pb.array, ph.array, meshed)
    d2h0, h2d0 = m.stats()["d2h_bytes"], m.stats()["h2d_bytes"]
    D.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        res = m.upload_and_mesh(resident, pb.array, ph.array, meshed)  # same three stages, overlapped
    D.sync()
    e2e_local = (time.perf_counter() - t0) * 1e3 / steps
    e2e_ms_step = D.reduce([e2e_local], "max")[0]
    s = m.stats()
    unscanned = int(s["upload_unscanned_chunks"])
    out = {"ms_per_step": ms_per_step, "value": tot_chunks / (ms_per_step / 1e3), "quads_per_s": tot_quads / (ms_per_step / 1e3),
           "total_chunks": tot_chunks, "total_quads": tot_quads, "total_launches": tot_launches, "total_worked": tot_worked,
           "counts": counts, "kernel_ms": {"derive": float(np.mean(derive_ms)), "classify": float(np.mean(classify_ms)), "mesh": float(np.mean(mesh_ms)),
                                           "source": "CUDA events of the ctx around each kernel of the last timed step", "warmup_steps_mean": warm},
           "e2e_ms_per_step": e2e_ms_step, "e2e_value": tot_chunks / (e2e_ms_step / 1e3), "e2e_serial_ms": e2e_serial_ms,
           "e2e_ms_per_rank": D.gather(e2e_local), "worked_per_rank": D.gather(counts["n_worked"]), "tasks_per_rank": D.gather(len(meshed)),
           "h2d_step": (s["h2d_bytes"] - h2d0) // steps, "d2h_step": (s["d2h_bytes"] - d2h0) // steps, "device_bytes": s["device_bytes"],
           "unscanned_last_upload": unscanned, "device_sink": None, "hinted": None, "parity": None, "cpu_stats": None, "variants": variants}
    if check:
        parity, st = compare_all_chunks(res, arm, meshed, cpu_threads)
        bad = D.reduce([parity["chunks_meshed"], parity["byte_compared_chunks"], parity["count_mismatches"], parity["byte_mismatches"]])
        parity.update(chunks_meshed=int(bad[0]), byte_compared_chunks=int(bad[1]), count_mismatches=int(bad[2]), byte_mismatches=int(bad[3]),
                      ranks=D.world)
        out["parity"], out["cpu_stats"] = parity, st
    if on_core is not None:
        on_core(out)
    # further end-to-end variants (they run new batches: `res` above is no longer valid from here on).  No collective inside them: a
    # rank that fails in one must not leave the others waiting; their local times are combined afterwards, in one reduction.
    sink = guarded("device sink", device_sink_steps, D, m, local, meshed, resident, pb, ph, steps)
    hinted = guarded("hinted upload", hinted_steps, D, m, meshed, resident, pb, ph, steps)
    fixed = guarded("fixed upload", fixed_upload_steps, D, m, meshed, resident, pb, ph, steps)
    legs = [sink, hinted, fixed]
    red = D.reduce([leg.get("local_ms", 0.0) if "error" not in leg else 0.0 for leg in legs], "max")
    bad = D.reduce([1.0 if "error" in leg else 0.0 for leg in legs])
    for leg, ms_, nbad in zip(legs, red, bad):
        if "error" not in leg:
            leg.pop("local_ms", None)
            if nbad:
                leg["error"] = f"failed on {int(nbad)} other rank(s)"
            else:
                leg["ms_per_step"], leg["value"], leg["unit"] = ms_, tot_chunks / (ms_ / 1e3), "chunks/s"
    out["device_sink"], out["hinted"], out["fixed_upload"] = sink, hinted, fixed
    guarded("close", m.close)
    return out


def device_buffer(D: Dist, nbytes: int, local: int):
    """A caller-owned device buffer (what a renderer's vertex / index buffer is): (raw pointer, keep-alive, read-back)."""
    if D.emulated:
        a = np.zeros(max(nbytes, 16), dtype=np.uint8)
        return a.ctypes.data, a, (lambda: a)
    t = D.torch.zeros(max(nbytes, 16), dtype=D.torch.uint8, device=f"cuda:{local}")
    D.sync()
    return t.data_ptr(), t, (lambda: t.cpu().numpy())


def device_sink_steps(D: Dist, m, local: int, meshed, resident, pb, ph, steps: int) -> dict:
    """SURVEY 8f row 2, the CUDA half of the zero-copy upload: the same world load as `e2e`, but the meshes stay on the device --
    hcm_upload_chunks + hcm_mesh_batch + hcm_mesh_counts (descriptors only) -> the host gives every mesh a range of ONE vertex and ONE
    index buffer the way MeshCluster::try_alloc does (client/include/client/mesh/MeshCluster.hpp:38-67: first fit in a virtual block,
    alignment sizeof(Vertex) / sizeof(Index); here a bump allocator over a fresh cluster) -> hcm_mesh_emit_into.  No vertex or index
    byte crosses the bus; the result is checked against the fetched (host) output of the same batch."""
    from hypercraft_b200 import mesher as M

    m.upload(resident, pb.array, ph.array)
    m.mesh_async(meshed)
    cnt = m.counts()
    nv, ni = m.n_vertices, m.n_indices
    vptr, vkeep, vread = device_buffer(D, nv * 8, local)
    iptr, ikeep, iread = device_buffer(D, ni * 2, local)

    def place(md):
        pl = np.zeros(len(md), dtype=M.MESH_PLACEMENT_DTYPE)
        pl["mesh"] = np.arange(len(md), dtype=np.uint32)
        pl["vertex_offset"][1:] = np.cumsum(md["vertex_count"][:-1], dtype=np.uint64)
        pl["index_offset"][1:] = np.cumsum(md["index_count"][:-1], dtype=np.uint64)
        return pl

    m.emit_into(place(cnt.meshes), vptr, nv, iptr, ni)
    d2h0, h2d0 = m.stats()["d2h_bytes"], m.stats()["h2d_bytes"]
    D.sync()
    t0 = time.perf_counter()
    sink_ms = []
    for _ in range(steps):
        m.upload(resident, pb.array, ph.array)
        m.mesh_async(meshed)
        cnt = m.counts()
        pl = place(cnt.meshes)
        m.emit_into(pl, vptr, nv, iptr, ni)
        sink_ms.append(m.stats()["sink_ms"])
    D.sync()
    local_ms = (time.perf_counter() - t0) * 1e3 / steps
    s = m.stats()
    # the device buffers hold what the host path returns for the same batch
    host = m.fetch()
    vb, ib = vread().view(np.uint64), iread().view(np.uint16)
    ok = len(host.meshes) == len(pl)
    for k in np.linspace(0, max(len(pl) - 1, 0), num=min(len(pl), 512), dtype=np.int64) if len(pl) else []:
        d = host.meshes[k]
        vo, io = int(pl[k]["vertex_offset"]), int(pl[k]["index_offset"])
        ok = ok and bool((vb[vo:vo + int(d["vertex_count"])] == host.vertices[int(d["vertex_offset"]):int(d["vertex_offset"]) + int(d["vertex_count"])]).all())
        ok = ok and bool((ib[io:io + int(d["index_count"])] == host.indices[int(d["index_offset"]):int(d["index_offset"]) + int(d["index_count"])]).all())
    del vkeep, ikeep
    return {"local_ms": local_ms, "h2d_bytes_per_step": int((s["h2d_bytes"] - h2d0) // steps),
            "d2h_bytes_per_step": int((s["d2h_bytes"] - d2h0) // steps), "d2h_vertex_index_bytes_per_step": 0, "scatter_kernel_ms": float(np.mean(sink_ms)),
            "device_vertex_bytes": int(nv * 8), "device_index_bytes": int(ni * 2), "meshes_placed": int(len(pl)), "device_bytes_equal_host_path": bool(ok),
            "call": "hcm_upload_chunks + hcm_mesh_batch + hcm_mesh_counts + host range allocation + hcm_mesh_emit_into (unpipelined)"}


def hinted_steps(D: Dist, m, meshed, resident, pb, ph, steps: int) -> dict:
    """NOT the headline (the hint is knowledge computed outside the timed region): the same world load with the per-chunk `uniform value`
    hint a HyperCraft build can pass -- its generator knows when it filled a chunk with one block (client/src/ChunkGenerateTask.cpp:21-39)
    -- so that neither the scan threads nor the copy engine touch the 64 KB of a uniform chunk.  Shows what the host-side scan costs."""
    b = pb.array
    uniform = (b[:, :64] == b[:, :1]).all(axis=1)
    idx = np.nonzero(uniform)[0]
    for i0 in range(0, len(idx), 1024):  # in slices: the candidates are two thirds of the world
        sel = idx[i0:i0 + 1024]
        uniform[sel] = (b[sel] == b[sel, :1]).all(axis=1)
    hint = np.where(uniform, b[:, 0].astype(np.int32), -2).astype(np.int32)  # HCM_HINT_NOT_UNIFORM for the rest
    m.upload(resident, pb.array, ph.array, uniform_hint=hint)
    m.mesh(meshed)
    D.sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        m.upload(resident, pb.array, ph.array, uniform_hint=hint)
        res = m.mesh(meshed)
    D.sync()
    return {"local_ms": (time.perf_counter() - t0) * 1e3 / steps, "uniform_chunks_hinted": int(uniform.sum()), "vertices": int(len(res.vertices)),
            "call": "hcm_upload_chunks_hinted + hcm_mesh_batch + hcm_mesh_fetch (unpipelined); compare with e2e.unpipelined_ms_per_step"}


def fixed_upload_steps(D: Dist, m, meshed, resident, pb, ph, steps: int) -> dict:
    """A/B of the upload balance, for information: the same hcm_upload_and_mesh steps as `e2e` with the round-1 upload (every chunk waits
    for its scan, HCM_UPLOAD_ADAPTIVE=0 behaviour) selected through hcm_set_option.  `e2e` stays the library's default mode."""
    m.set_option("upload_adaptive", 0)
    try:
        m.upload_and_mesh(resident, pb.array, ph.array, meshed)
        D.sync()
        t0 = time.perf_counter()
        for _ in range(steps):
            m.upload_and_mesh(resident, pb.array, ph.array, meshed)
        D.sync()
        ms = (time.perf_counter() - t0) * 1e3 / steps
    finally:
        m.set_option("upload_adaptive", 1)
    return {"local_ms": ms, "call": "hcm_upload_and_mesh with every chunk scanned by the host first (round-1 upload mode)"}


def config4_block(D: Dist, args, rank: int, world: int, local: int, threads: int) -> dict:
    """BASELINE.json configs[3]: ONE r = 32 load sphere, its task list cut into `world` x-slabs, each GPU holding its slab dilated by one
    chunk (hypercraft_b200/shard.py; no collective on the data path).  The cut balances ESTIMATED WORK, not task counts: most chunks of
    a sphere are all air or all stone and never reach the mesh kernel, so every rank generates a thin sample of the world (its share
    of the x-planes, every 4th z-column), counts the non-uniform chunks per plane, and the per-plane estimates are summed over the
    ranks -- a few hundred bytes of host bookkeeping.  Reference: misc/load_radius.cpp:17-35 (the list), SURVEY 8e (the partition)."""
    from hypercraft_b200 import shard as sh
    from hypercraft_b200 import worldgen

    radius = args.config4_radius
    t0 = time.time()
    meshed_all, _ = worldgen.load_radius_positions(radius)
    arm, provide, _ = make_cpu_arm(threads)
    # -- work estimate per x-plane from a sample of the terrain -------------------------------------------------------------------
    x_min = int(meshed_all[:, 0].min())
    n_planes = int(meshed_all[:, 0].max()) - x_min + 1
    zstride = 4
    mine = ((meshed_all[:, 0].astype(np.int32) - x_min) % world == rank) & ((meshed_all[:, 2].astype(np.int32) - x_min) % zstride == 0)
    sample = meshed_all[mine]
    provide(sample)
    est = np.zeros(n_planes)
    if len(sample):
        sb = np.empty((len(sample), 32768), dtype=np.uint16)
        shh = np.empty((len(sample), 1024), dtype=np.int8)
        if isinstance(arm, PortArm):
            for i, p in enumerate(sample):
                sb[i], shh[i] = arm.chunks[tuple(int(c) for c in p)]
        else:
            arm.get_chunks(sample, sb, shh)
        nonuniform = (sb != sb[:, :1]).any(axis=1)
        np.add.at(est, sample[:, 0].astype(np.int32) - x_min, nonuniform * float(zstride))
        del sb, shh
    est = np.array(D.reduce(est.tolist()))
    tasks_per_plane = np.bincount(meshed_all[:, 0].astype(np.int32) - x_min, minlength=n_planes).astype(np.float64)
    # SM-time per unit (measured at r = 20): ~41 us per worked chunk in mesh_kernel, ~4.7 us per resident chunk in derive_kernel
    frac_worked = np.divide(est, tasks_per_plane, out=np.zeros_like(est), where=tasks_per_plane > 0)
    weights = 4.7 + 41.0 * frac_worked[meshed_all[:, 0].astype(np.int32) - x_min]
    bounds = sh.slab_bounds(meshed_all, world, weights)
    meshed = sh.shard_tasks(meshed_all, world, rank, weights)
    resident = sh.resident_for(meshed)
    provide(resident)
    pb, ph = fill_pinned(arm, resident)
    origin, extent = sh.window_for(resident)
    setup_s = time.time() - t0
    r = measure(D, local, meshed, resident, pb, ph, origin, extent, args.steps, args.warmup, arm, threads, not args.no_cpu_baseline)
    pb.free(), ph.free()
    worked = r["worked_per_rank"]
    return {"workload": workload_name(radius, True), "radius": radius, "scaling": "strong", "chunks_meshed": int(r["total_chunks"]),
            "value": r["value"], "unit": "chunks/s", "ms_per_step": r["ms_per_step"], "quads_per_s": r["quads_per_s"],
            "e2e": {"value": r["e2e_value"], "unit": "chunks/s", "ms_per_step": r["e2e_ms_per_step"], "ms_per_rank": r["e2e_ms_per_rank"],
                    "h2d_bytes_per_step_rank0": int(r["h2d_step"]), "d2h_bytes_per_step_rank0": int(r["d2h_step"]),
                    "chunks_sent_unscanned_last_step_rank0": r["unscanned_last_upload"]},
            "e2e_device_sink": r["device_sink"],
            "kernel_ms_rank0": r["kernel_ms"], "slab_x_bounds": [[int(a), int(b)] for a, b in bounds],
            "tasks_per_rank": [int(v) for v in r["tasks_per_rank"]], "worked_chunks_per_rank": [int(v) for v in worked],
            "worked_imbalance_max_over_mean": float(max(worked) / max(np.mean(worked), 1.0)),
            "balance": "x-slabs balanced on estimated work per x-plane: 41 us x (sampled non-uniform chunks) + 4.7 us x tasks; the sample is every "
                       "4th z-column of the terrain, generated by the ranks in turn and summed (host bookkeeping, no data-path collective)",
            "halo": "each slab's boundary chunk planes are uploaded from the host to both neighbours (no peer copies): the chunks live in host Chunk "
                    "objects anyway and a world load moves every chunk over PCIe once per holder",
            "parity": r["parity"], "setup_seconds": setup_s}


def sharded_dispatcher_check(emulated: bool) -> dict:
    """N > 1 only, rank 0: the C++ host mirror for several GPUs -- hcm_host::ShardedMeshDispatcher, one hcm_ctx per device over x-slabs,
    FlushParallel with one host thread per GPU, edits across the cut -- run through the C ABI on devices 0 and 1 of this box and compared
    mesh for mesh with the CPU oracle (tests/cpp/test_sharded_gpu.cpp).  A correctness check riding along with the scaling run, not a
    measurement."""
    import hcoracle
    from hypercraft_b200 import mesher as M

    t0 = time.time()
    try:
        exe = os.path.join(ROOT, "tests", "cpp", "test_sharded_gpu")
        libdir, libfile = os.path.split(M.LIB_PATH)
        subprocess.run(["g++", "-std=c++20", "-O2", "-pthread", "-o", exe, exe + ".cpp", f"-L{libdir}", "-l" + libfile[3:-3], "-ldl", f"-Wl,-rpath,{libdir}"],
                       check=True, capture_output=True, timeout=300)
        hcoracle.build(ref=False, port=True)
        env = dict(os.environ, CUSIM_DEVICES="2") if emulated else dict(os.environ)
        r = subprocess.run([exe, os.path.join(ROOT, "hypercraft_b200", "data", "registry.bin"), hcoracle.PORT_SO], capture_output=True, text=True, timeout=300, env=env)
        lines = r.stdout.strip().splitlines()
        return {"devices": 2, "returncode": r.returncode, "result": lines[-1] if lines else r.stderr[-300:], "detail": lines[:-1], "seconds": time.time() - t0,
                "what": "tests/cpp/test_sharded_gpu.cpp: ShardedMeshDispatcher world load (FlushParallel) + edits across the cut on two devices, every mesh vs the oracle"}
    except Exception as e:  # a check that cannot run must not take the scaling numbers down
        return {"devices": 2, "error": f"{type(e).__name__}: {e}", "seconds": time.time() - t0}


def secondary_configs(local: int, threads: int) -> dict:
    """BASELINE.json configs[2] (worst-case face density) and configs[4] (edit-driven remesh), small and bounded; each result is
    compared with the CPU arm chunk for chunk.  See tools/bench_configs.py."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_configs

    out = {}
    try:
        out["config3_worst_case"] = bench_configs.config3(local, threads)
        out["config5_edits_on_device"] = bench_configs.config5_device(local, threads)
    except Exception as e:  # secondary figures never take the headline down
        out["error"] = f"{type(e).__name__}: {e}"
    return out


def main():
    args = parse_args()
    rank, world, local = dist_env()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch

    emulated = args.dry_run_emulated
    if emulated:  # developer dry run of this script's logic on the CPU emulation build (tests/cusim): never a measurement
        from hypercraft_b200 import mesher as M

        sim_dir = os.path.join(ROOT, "tests", "cusim")
        subprocess.run(["make", "-C", sim_dir], check=True, stdout=subprocess.DEVNULL)
        M.LIB_PATH, M._lib = os.path.join(sim_dir, "_build", "libhcm_sim.so"), None
        local = 0
    elif not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the mesher has no CPU path (use --impl reference for the CPU baseline)")
    else:
        torch.cuda.set_device(local)
    D = Dist(world, local, emulated)

    from hypercraft_b200 import shard as sh
    from hypercraft_b200 import worldgen

    radius = args.radius
    threads = max(1, (os.cpu_count() or 1) // world)
    os.environ.setdefault("HCM_HOST_THREADS", str(max(1, min(15, threads - 1))))  # upload scan threads of this rank's hcm_ctx
    affinity = None
    if world > 1:  # ranks share one host: each keeps to its own cores (the scan threads and the CPU arm inherit the mask)
        affinity = pin_rank_to_cores(local, world, emulated)
        if affinity and affinity.get("cores"):
            threads = max(1, len(affinity["cores"]))
            os.environ["HCM_HOST_THREADS"] = str(max(1, min(15, threads - 1)))
    check = not args.no_cpu_baseline
    sampler = ClockSampler(local)
    sampler.start()
    wd = Watchdog(rank)

    if args.strong:  # config 4 as the headline
        args.config4_radius = radius
        wd.arm("config4 as the headline", 2400)
        c4 = config4_block(D, args, rank, world, local, threads)
        wd.deadline = None
        clocks = sampler.stop()
        if rank == 0:
            print(json.dumps({"metric": "chunks meshed/sec", "value": c4["value"], "unit": "chunks/s", "n_gpus": world, "steps": args.steps,
                              "warmup": max(args.warmup, 3), "ms_per_step": c4["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                              "vs_baseline": None, "dtype": "u8/u16 integer", "data": f"synthetic: DefaultTerrain(seed {SEED}) or repo terrain, see config4",
                              "config": {"workload": c4["workload"], "radius": radius}, "quads_per_s": c4["quads_per_s"],
                              "e2e": c4["e2e"], "config4": c4, "parity": c4["parity"], "clocks": clocks, **({"emulated": True} if emulated else {})}))
        D.close()
        return

    centre = (3 * radius * rank, 0, 0)
    t_setup = time.time()
    wd.arm("set-up (terrain generation)", 900)
    meshed, resident = worldgen.load_radius_positions(radius)
    c = np.asarray(centre, dtype=np.int16)
    meshed, resident = (meshed + c).astype(np.int16), (resident + c).astype(np.int16)
    arm, provide, data_label = make_cpu_arm(threads)
    provide(resident)
    pb, ph = fill_pinned(arm, resident)
    ext = 2 * radius + 3
    origin, extent = (centre[0] - radius - 1, -radius - 1, -radius - 1), (ext, ext, ext)
    setup_s = time.time() - t_setup

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
    traffic = None
    if radius == 20:  # the committed ncu launch list is of this workload; other radii have no measured traffic figure
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json")))
        except Exception:
            pass

    # everything the line is built from; optional blocks fill their slots as they complete (the watchdog prints what is there)
    S = {"r": None, "clocks": None, "cpu_baseline": None, "config4": None, "sharded": None, "secondary": None}

    def build_line() -> dict:
        r = S["r"]
        counts = r["counts"]
        out = {
            "metric": "chunks meshed/sec", "value": r["value"], "unit": "chunks/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u16 integer",
            "data": data_label,
            "config": {"workload": workload_name(radius), "radius": radius,
                       "chunks_meshed_per_gpu": int(len(meshed)), "chunks_resident_per_gpu": int(len(resident)),
                       "tasks_through_full_kernel_per_gpu": int(counts["n_worked"]),
                       "parallelism": f"{world} independent spheres (no collective)",
                       "rank0_cpu_affinity": ({"n_cores": len(affinity["cores"]), "how": affinity["how"]} if affinity and affinity.get("cores") else None),
                       "l2": f"inputs ({r['device_bytes'] / 1e9:.1f} GB of device-resident world per GPU) larger than L2; no flush needed",
                       "setup_seconds": setup_s},
            "quads_per_s": r["quads_per_s"], "quads_per_step": r["total_quads"], "vertices_per_step_per_gpu": int(counts["n_vertices"]),
            "e2e": {"value": r["e2e_value"], "unit": "chunks/s", "ms_per_step": r["e2e_ms_per_step"], "h2d_bytes_per_step": int(r["h2d_step"]),
                    "d2h_bytes_per_step": int(r["d2h_step"]), "call": "hcm_upload_and_mesh", "unpipelined_ms_per_step": r["e2e_serial_ms"],
                    "ms_per_rank": r["e2e_ms_per_rank"], "chunks_sent_unscanned_last_step": r["unscanned_last_upload"],
                    "upload": "scan threads and copy engine share the work items: what the engine takes first crosses the bus unscanned "
                              "(derive_kernel finds uniform chunks on the device as well)"},
            "e2e_device_sink": r.get("device_sink"), "e2e_with_generator_hints": r.get("hinted"), "e2e_round1_upload_mode": r.get("fixed_upload"),
            "gpu_launches": int(r["total_launches"]),
            "kernel_ms": r["kernel_ms"], "kernel_variants": r.get("variants"),
            "roofline": roofline_block(len(meshed), int(counts["n_worked"]), len(resident), int(counts["n_vertices"]), int(counts["n_meshes"]),
                                       r["kernel_ms"], r["ms_per_step"], peak_gbs, peak_src, traffic),
            "cpu_baseline": S["cpu_baseline"], "parity": r["parity"], "clocks": S["clocks"],
        }
        if S["config4"] is not None:
            out["config4"] = S["config4"]
        if S["sharded"] is not None:
            out["sharded_dispatcher"] = S["sharded"]
        if S["secondary"] is not None:
            out["secondary_configs"] = S["secondary"]
        if emulated:
            out = {"emulated": True, "note": "DRY RUN on the CPU emulation build (tests/cusim): exercises bench.py's logic, measures nothing", **out}
        return out

    def core_done(core: dict):
        # headline numbers (value, e2e, parity) exist: from here on a line is printed whatever happens to the optional blocks
        S["r"] = core
        S["clocks"] = sampler.stop()  # sampled from the first warm-up step to the last end-to-end step
        if rank == 0 and world == 1 and check and core["cpu_stats"]:  # the CPU baseline is reported at N = 1 only; the same call was the parity gate's checker
            st = core["cpu_stats"]
            S["cpu_baseline"] = {"value": st["chunks"] / st["seconds"], "unit": "chunks/s", "cores": threads, "kind": getattr(arm, "kind", "reference"),
                                 "quads_per_s": st["vertices"] / 4 / st["seconds"], "cpu_model": cpu_model(),
                                 "sample": f"{len(meshed)} of {len(meshed)} chunks (the whole r={radius} task list, {st['seconds']:.1f} s), "
                                           + getattr(arm, "what", "unmodified ChunkTaskRunner<kMesh>::Run, one runner per thread")}
        wd.line_fn = build_line
        wd.arm("optional end-to-end variants (device sink, hinted upload, round-1 upload mode)", 420)

    wd.arm("headline measurement (device-timed steps, end-to-end steps, parity gate)", 1200)
    r = measure(D, local, meshed, resident, pb, ph, origin, extent, args.steps, args.warmup, arm, threads, check, on_core=core_done)
    S["r"] = r
    pb.free(), ph.free()

    if S["cpu_baseline"] is not None:
        wd.arm("cpu_baseline at the reference client's worker count", 300)
        S["cpu_baseline"]["reference_pool_threads"] = guarded("pool baseline", pool_baseline, arm, meshed[:: max(1, len(meshed) // 8192)], threads)
    del arm

    if world > 1 and args.config4_radius > 0:
        wd.arm("config4 (one r = %d sphere in x-slabs)" % args.config4_radius, 1200)
        S["config4"] = config4_block(D, args, rank, world, local, threads)
    if world == 1 and rank == 0 and not args.no_secondary and check:
        wd.arm("secondary_configs (configs 3 and 5)", 600)
        S["secondary"] = secondary_configs(local, threads)

    wd.arm("last barrier", 300)
    D.barrier()  # the last collective: what follows is rank 0's alone
    if world > 1 and rank == 0:  # after every collective: the other ranks are done
        wd.arm("sharded dispatcher check", 700)
        S["sharded"] = sharded_dispatcher_check(emulated)
    wd.deadline = None
    if rank == 0:
        wd.emit(build_line())
    guarded("close", D.close)


if __name__ == "__main__":
    main()
