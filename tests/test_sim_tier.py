"""CPU tier of the kernel tests: the `gpu`-marked parity suite run against tests/cusim -- the CUDA sources of
hypercraft_b200/csrc compiled with g++ on top of a CPU emulation of the CUDA execution model (fibres for threads, rendez-vous
for warp collectives and __syncthreads, mbarrier + TMA completing at issue).  TEST INFRASTRUCTURE: it proves arithmetic,
indexing, barrier structure and the host-side call sequences on a machine without a GPU; it proves nothing about timing or the
hardware's memory model, and the product never loads it.  `pytest -m gpu` on a B200 remains the gate.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env, *pytest_args):
    env = dict(os.environ, HCM_TEST_SIM="1", **extra_env)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), os.path.join(ROOT, "tests", "test_zz_kernel_variants.py"),
                        "-m", "gpu", "-x", "-q", "-p", "no:cacheprovider", *pytest_args], capture_output=True, text=True, env=env, cwd=ROOT)
    tail = "\n".join((r.stdout + r.stderr).splitlines()[-25:])
    assert r.returncode == 0, tail
    return tail


def test_gpu_parity_suite_under_cpu_emulation():
    """Every GPU parity test (oracle bytes, reference digests, pipelined uploads, edits, device sink, the C++ host mirror on one and
    on two emulated devices) passes on the emulation."""
    tail = _run({}, "-n", "4")  # four pytest workers: most tests launch one CTA
    assert " passed" in tail and "failed" not in tail


def test_kernels_under_the_reversed_schedule():
    """The emulation runs one warp at a time, lowest first; CUSIM_REVERSE visits warps and lanes highest first.  A missing barrier
    between a producer and a consumer warp gives stale data under one of the two orders."""
    _run({"CUSIM_REVERSE": "1"}, "-n", "4", "-k", "mesh_matches_oracle or light_stage or set_blocks_on_device or device_sink or forced_variants or ab_rounds")


def test_kernel_tests_with_the_variant_kernels_forced():
    """HCM_DERIVE_VARIANT=3 HCM_MESH_VARIANT=2: every context of the parity suite's kernel tests launches variant forms of derive_kernel
    and mesh_kernel (include/hcm.h, "Kernel variants") -- the same oracle bytes and reference digests must come out.  (The other forms
    are forced test by test in tests/test_zz_kernel_variants.py, which the first test of this file runs on the emulation.)"""
    tail = _run({"HCM_DERIVE_VARIANT": "3", "HCM_MESH_VARIANT": "2"}, "-n", "4", "--deselect",
                "tests/test_zz_kernel_variants.py::test_ab_rounds_decide_and_never_change_a_byte", "--deselect",
                "tests/test_zz_kernel_variants.py::test_a_differing_hash_rejects_the_variant_for_good", "--deselect",
                "tests/test_zz_kernel_variants.py::test_tuning_can_be_switched_off", "-k",
                "mesh_matches_oracle or light_stage or terrain_goldens or whole_load_sphere or pipelined_world_load or set_blocks_matches_the_reference "
                "or mixed_neighbourhood or batch_of_many or output_arena_grows or edit_driven")  # the kernels' tests; host-only tests do not care
    assert " passed" in tail and "failed" not in tail


def test_upload_paths_under_a_busy_copy_engine():
    """CUSIM_EVENT_BUSY makes cudaEventQuery answer 'not ready' a few times first, so the uploading thread of a threaded upload waits
    for the scan threads instead of taking every work item itself: mixed scanned / unscanned blocks."""
    _run({"CUSIM_EVENT_BUSY": "40"}, "-k", "pipelined_world_load or hinted_upload")


def test_every_large_enough_call_as_an_ab_round():
    """HCM_TUNE_MIN_CHUNKS=1 HCM_TUNE_MIN_TASKS=1: every derive launch and every mesh batch of a fresh context is an A/B round (baseline,
    every variant, baseline, hashes compared) until the context has decided -- through uploads, pipelined world loads, edits, the device
    sink, halo copies and re-derives.  Results must not change by a byte."""
    _run({"HCM_TUNE_MIN_CHUNKS": "1", "HCM_TUNE_MIN_TASKS": "1"}, "-n", "4", "-k",
         "(mesh_matches_oracle and (mix or flow or checker)) or pipelined_world_load or set_blocks_on_device or halo_replication or translation_invariance")


def _bench_dry_run(n_ranks, *args):
    import json

    cmd = [sys.executable]
    if n_ranks > 1:
        cmd += ["-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n_ranks}", "--master-addr", "127.0.0.1", "--master-port", "29533"]
    cmd += [os.path.join(ROOT, "bench.py"), "--gpus", str(n_ranks), "--dry-run-emulated", "--steps", "1", "--warmup", "1", *args]
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=900)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert r.returncode == 0 and len(lines) == 1, r.stdout[-2000:] + r.stderr[-3000:]
    return json.loads(lines[0])


def test_bench_logic_dry_run_on_the_emulation():
    """bench.py's own Python -- workload set-up, the three timed regions, per-chunk parity digests, roofline and CPU-baseline blocks --
    exercised end to end at a tiny radius on the emulation (`--dry-run-emulated`: the line is flagged and its numbers mean nothing)."""
    d = _bench_dry_run(1, "--radius", "3", "--no-secondary")
    assert d["emulated"] is True and d["n_gpus"] == 1 and d["gpu_launches"] == 3
    p = d["parity"]
    assert p["chunks_meshed"] == p["byte_compared_chunks"] == d["config"]["chunks_meshed_per_gpu"] and p["byte_mismatches"] == 0 and p["count_mismatches"] == 0
    rf = d["roofline"]
    assert rf["units_per_launch"] == d["config"]["tasks_through_full_kernel_per_gpu"] and rf["derive_kernel"]["units_per_launch"] == d["config"]["chunks_resident_per_gpu"]
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12 and rf["step_frac"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["cpu_baseline"]["cores"] >= 1


def test_bench_two_ranks_dry_run_on_the_emulation():
    """--gpus 2 under torchrun (gloo in the dry run): every rank's chunks go through the parity gate (counts summed over the ranks) and
    the config-4 block cuts ONE sphere into two x-slabs whose union is the whole task list."""
    d = _bench_dry_run(2, "--radius", "3", "--config4-radius", "4")
    assert d["emulated"] is True and d["n_gpus"] == 2 and d["cpu_baseline"] is None
    p = d["parity"]
    assert p["ranks"] == 2 and p["byte_compared_chunks"] == p["chunks_meshed"] == 2 * d["config"]["chunks_meshed_per_gpu"] and p["byte_mismatches"] == 0
    c4 = d["config4"]
    assert c4["scaling"] == "strong" and sum(c4["tasks_per_rank"]) == c4["chunks_meshed"] == c4["parity"]["byte_compared_chunks"]
    assert c4["parity"]["byte_mismatches"] == 0 and len(c4["slab_x_bounds"]) == 2 and c4["slab_x_bounds"][0][1] == c4["slab_x_bounds"][1][0]
    assert d["sharded_dispatcher"]["returncode"] == 0 and d["sharded_dispatcher"]["result"].startswith("PASS"), d["sharded_dispatcher"]
