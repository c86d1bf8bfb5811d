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
all ranks / max-over-ranks time.

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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--radius", type=int, default=20)
    ap.add_argument("--strong", action="store_true",
                    help="BASELINE config 4: ONE load sphere cut into x-slabs of equal task count, one slab (dilated by a chunk) per GPU")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU work of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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


def build_world(radius: int, centre, threads: int, shard=None):
    """(meshed positions, resident positions, pinned blocks, pinned heights, data label, CPU arm: RefWorld or PortArm).
    shard = (world, rank): only this rank's x-slab of the task list and its one-chunk dilation (hypercraft_b200/shard.py)."""
    import hcoracle
    from hypercraft_b200 import mesher, worldgen

    meshed, resident = worldgen.load_radius_positions(radius)
    c = np.asarray(centre, dtype=np.int16)
    meshed = (meshed + c).astype(np.int16)
    resident = (resident + c).astype(np.int16)
    if shard is not None:
        from hypercraft_b200 import shard as sh

        # equal task counts: measured better (2.40 ms at r=32 on 8 GPUs) than shard.surface_weights' y-band estimate (2.65 ms)
        meshed = sh.shard_tasks(meshed, shard[0], shard[1])
        resident = sh.resident_for(meshed)
    pb = mesher.PinnedArray((len(resident), 32768), np.uint16)
    ph = mesher.PinnedArray((len(resident), 1024), np.int8)
    ref = None
    if use_reference_build():
        ref = hcoracle.RefWorld()
        ref.generate(resident, seed=SEED, threads=threads)
        ref.get_chunks(resident, pb.array, ph.array)
        label = f"synthetic: reference DefaultTerrain(seed {SEED}) run on this host via oracle/_ref (no dataset)"
    else:
        for i, p in enumerate(resident):
            pb.array[i], ph.array[i] = worldgen.synth_terrain_chunk(SEED, p)
        ref = PortArm({tuple(int(c) for c in p): (pb.array[i], ph.array[i]) for i, p in enumerate(resident)})
        label = "synthetic: repo height-field terrain (oracle/_ref terrain generator not available)"
    return meshed, resident, pb, ph, label, ref


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
    """SURVEY 8d, per mesh_kernel launch: the canonical figure roofline.achieved uses (34^3 halo blocks 2 B + light 1 B per chunk, 8 B
    per vertex written) and the two secondary ones -- "full": what the reference task really touches (62^3 blocks + 3 x 62^2 heights,
    light re-derived from raw blocks); "min": every chunk read once from HBM with perfect halo reuse, vertices + indices + mesh infos."""
    return {"canonical": n_chunks * H_BYTES + 8 * n_vertices,
            "full": n_chunks * (62 ** 3 * 2 + 62 * 62 * 3) + 8 * n_vertices,
            "min": n_chunks * 66560 + 11 * n_vertices + 40 * n_meshes}


def roofline_secondary(algo: dict, kernel_ms: float, peak_gbs: float) -> dict:
    """The two secondary byte models of SURVEY 8d against the same launch time and peak as roofline.achieved."""
    return {"secondary": {k: {"algorithmic_bytes_per_launch": int(algo[k]), "achieved": algo[k] / (kernel_ms / 1e3) / 1e9,
                              "frac": algo[k] / (kernel_ms / 1e3) / 1e9 / peak_gbs} for k in ("full", "min")},
            "secondary_formulas": {"full": "62^3*2 + 62*62*3 + 8*V per chunk (what the reference task touches)",
                                   "min": "66560 + 11*V + 40*M per chunk (each chunk read once, vertices + indices + mesh infos)"}}


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


def use_reference_build() -> bool:
    """oracle/_ref when it was compiled (it travels to the GPU box); HCM_BENCH_FORCE_PORT=1 selects the port (test knob)."""
    import hcoracle

    return hcoracle.have_ref() and not os.environ.get("HCM_BENCH_FORCE_PORT")


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path (unmodified ChunkTaskRunner<kMesh>::Run from oracle/_ref, all host
    threads); the oracle port (cpu_baseline.kind "port") where oracle/_ref was not built."""
    if rank != 0:
        return
    import hcoracle
    from hypercraft_b200 import worldgen

    threads = os.cpu_count() or 1
    radius = args.radius
    meshed, resident = worldgen.load_radius_positions(radius)
    if use_reference_build():
        ref, kind, data = hcoracle.RefWorld(), "reference", f"synthetic: reference DefaultTerrain(seed {SEED})"

        def provide(need):  # neighbours generated on demand
            ref.generate(need, seed=SEED, threads=threads)
    else:  # the reference did not compile here: its restatement, on the terrain the b200 arm falls back to as well
        ref, kind, data = PortArm({}), "port", "synthetic: repo height-field terrain (oracle/_ref terrain generator not available)"

        def provide(need):
            for q in need:
                key = tuple(int(c) for c in q)
                if key not in ref.chunks:
                    ref.chunks[key] = worldgen.synth_terrain_chunk(SEED, q)
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


def main():
    args = parse_args()
    rank, world, local = dist_env()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the mesher has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from hypercraft_b200 import mesher as M

    radius = args.radius
    threads = max(1, (os.cpu_count() or 1) // world)
    os.environ.setdefault("HCM_HOST_THREADS", str(max(1, min(15, threads - 1))))  # upload scan threads of this rank's hcm_ctx
    centre = (0, 0, 0) if args.strong else (3 * radius * rank, 0, 0)
    t_setup = time.time()
    meshed, resident, pb, ph, data_label, ref = build_world(radius, centre, threads, (world, rank) if args.strong else None)
    if args.strong:
        from hypercraft_b200 import shard as sh

        origin, extent = sh.window_for(resident)
    else:
        ext = 2 * radius + 3
        origin, extent = (centre[0] - radius - 1, -radius - 1, -radius - 1), (ext, ext, ext)
    m = M.Mesher(origin=origin, extent=extent, device=local, max_chunks=len(resident), max_tasks=len(meshed))
    m.upload(resident, pb.array, ph.array)
    m.sync()
    setup_s = time.time() - t_setup

    stream = torch.cuda.ExternalStream(m.stream_ptr(), device=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        m.rederive()
        m.mesh_async(meshed)
        return m.wait()

    sampler = ClockSampler(local)
    sampler.start()
    # ---- warm-up (also sizes the output arenas) -------------------------------------------------------------------
    counts = None
    for _ in range(max(args.warmup, 3)):
        counts = step_resident()
    n_quads = counts["n_vertices"] // 4
    algo = algorithmic_bytes(len(meshed), int(counts["n_vertices"]), int(counts["n_meshes"]))
    algo_bytes = algo["canonical"]

    # ---- timed: K steps, device timed on the ctx stream, max over ranks --------------------------------------------
    barrier()
    launches0 = m.stats()["kernel_launches"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    mesh_ms, derive_ms, classify_ms = [], [], []
    e0.record(stream)
    for _ in range(args.steps):
        step_resident()
        s = m.stats()
        mesh_ms.append(s["mesh_ms"]), derive_ms.append(s["derive_ms"]), classify_ms.append(s["classify_ms"])
    e1.record(stream)
    barrier()
    launches = m.stats()["kernel_launches"] - launches0
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    tot = torch.tensor([len(meshed), int(n_quads), int(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_per_step = float(t.item()) / args.steps
    total_chunks, total_quads, total_launches = (float(x) for x in tot.tolist())
    value = total_chunks / (ms_per_step / 1e3)

    # ---- end to end: host buffers in, host meshes out, every step ---------------------------------------------------
    res = None
    barrier()
    for it in range(3):  # the three calls one after the other (first pass: warm-up of the pinned result buffers)
        if it == 1:
            t0 = time.perf_counter()
        m.upload(resident, pb.array, ph.array)
        res = m.mesh(meshed)
    e2e_serial_ms = (time.perf_counter() - t0) * 1e3 / 2
    res = m.upload_and_mesh(resident, pb.array, ph.array, meshed)
    d2h0, h2d0 = m.stats()["d2h_bytes"], m.stats()["h2d_bytes"]
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = m.upload_and_mesh(resident, pb.array, ph.array, meshed)  # same three stages, overlapped
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms_step = float(t.item()) / args.steps
    s = m.stats()
    h2d_step = (s["h2d_bytes"] - h2d0) // args.steps
    d2h_step = (s["d2h_bytes"] - d2h0) // args.steps
    e2e_value = total_chunks / (e2e_ms_step / 1e3)
    clocks = sampler.stop()  # sampled from the first warm-up step to the last end-to-end step

    # ---- parity gate + CPU baseline on rank 0 (bounded sample of the same workload) -----------------------------------
    cpu_baseline, parity = None, None
    if rank == 0 and ref is not None and not args.no_cpu_baseline:
        import hcoracle

        cpu_threads = os.cpu_count() or 1
        probe = meshed[:: max(1, len(meshed) // 128)][:128]
        st, _ = ref.mesh_batch(probe, threads=cpu_threads)
        rate = st["chunks"] / max(st["seconds"], 1e-6)
        n_s = int(min(len(meshed), max(128, rate * args.cpu_seconds)))
        stride = max(1, len(meshed) // n_s)
        idx = np.arange(0, len(meshed), stride)[:n_s]
        st, per = ref.mesh_batch(meshed[idx], threads=cpu_threads, per_chunk=True)
        cpu_baseline = {"value": st["chunks"] / st["seconds"], "unit": "chunks/s", "cores": cpu_threads, "kind": getattr(ref, "kind", "reference"),
                        "quads_per_s": st["vertices"] / 4 / st["seconds"], "cpu_model": cpu_model(),
                        "sample": f"{len(idx)} of {len(meshed)} chunks (every {stride}th of the r={radius} task list), "
                                  + getattr(ref, "what", "unmodified ChunkTaskRunner<kMesh>::Run, one runner per thread")}
        # the reference's own worker count (Application.cpp:248: hardware_concurrency() * 3 / 4), same sample
        cpu_baseline["reference_pool_threads"] = pool_baseline(ref, meshed[idx], cpu_threads)
        # parity: mesh/vertex/index counts of every sampled chunk + full bytes of the heaviest ones against the reference
        got_m = res.tasks["n_meshes"][idx].astype(np.uint64)
        first = res.tasks["first_mesh"][idx]
        gv = np.array([int(res.meshes["vertex_count"][f:f + n].sum()) for f, n in zip(first, got_m)], dtype=np.uint64)
        gi = np.array([int(res.meshes["index_count"][f:f + n].sum()) for f, n in zip(first, got_m)], dtype=np.uint64)
        count_bad = int(((got_m != per[:, 0]) | (gv != per[:, 1]) | (gi != per[:, 2])).sum())
        heavy = idx[np.argsort(-per[:, 1].astype(np.int64))[:48]]
        byte_bad = 0
        for t_i in heavy:
            a, b = res.serialise(int(t_i)), ref.mesh(meshed[t_i])
            byte_bad += not (a.size == b.size and (a == b).all())
        parity = {"sampled_chunks": int(len(idx)), "count_mismatches": count_bad, "byte_compared_chunks": int(len(heavy)), "byte_mismatches": int(byte_bad)}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_gbs, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
        traffic = None
        if radius == 20:  # the ncu capture is of this workload; other radii have no measured traffic figure
            try:
                traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json"))).get("mesh_kernel_dram_bytes_per_launch")
            except Exception:
                pass
        k_ms = float(np.mean(mesh_ms))
        achieved = algo_bytes / (k_ms / 1e3) / 1e9
        out = {
            "metric": "chunks meshed/sec", "value": value, "unit": "chunks/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "u8/u16 integer",
            "data": data_label,
            "config": {"workload": workload_name(radius, args.strong), "radius": radius,
                       "chunks_meshed_per_gpu": int(len(meshed)), "chunks_resident_per_gpu": int(len(resident)),
                       "tasks_through_full_kernel_per_gpu": int(counts["n_worked"]),
                       "parallelism": (f"{world} x-slabs of one sphere, each dilated by one chunk (no collective); per-GPU figures are rank 0's"
                                       if args.strong else f"{world} independent spheres (no collective)"),
                       "l2": f"inputs ({m.stats()['device_bytes'] / 1e9:.1f} GB of device-resident world per GPU) larger than L2; no flush needed",
                       "setup_seconds": setup_s},
            "quads_per_s": total_quads / (ms_per_step / 1e3), "quads_per_step": total_quads, "vertices_per_step_per_gpu": int(counts["n_vertices"]),
            "e2e": {"value": e2e_value, "unit": "chunks/s", "ms_per_step": e2e_ms_step, "h2d_bytes_per_step": int(h2d_step),
                    "d2h_bytes_per_step": int(d2h_step), "call": "hcm_upload_and_mesh", "unpipelined_ms_per_step": e2e_serial_ms},
            "gpu_launches": int(total_launches),
            "kernel_ms": {"derive": float(np.mean(derive_ms)), "classify": float(np.mean(classify_ms)), "mesh": k_ms},
            "roofline": {"bound": "hbm", "kernel": "hcm::mesh_kernel", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(algo_bytes), "formula": "sum over chunks of 34^3*3 + 8*vertices (SURVEY 8d)",
                         **roofline_secondary(algo, k_ms, peak_gbs)},
            "cpu_baseline": cpu_baseline, "parity": parity, "clocks": clocks,
        }
        print(json.dumps(out))
    m.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
