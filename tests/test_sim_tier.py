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
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu", "-x", "-q", "-p", "no:cacheprovider",
                        *pytest_args], capture_output=True, text=True, env=env, cwd=ROOT)
    tail = "\n".join((r.stdout + r.stderr).splitlines()[-25:])
    assert r.returncode == 0, tail
    return tail


def test_gpu_parity_suite_under_cpu_emulation():
    """Every GPU parity test (oracle bytes, reference digests, pipelined uploads, edits, device sink, the C++ host mirror on one and
    on two emulated devices) passes on the emulation."""
    tail = _run({})
    assert " passed" in tail and "failed" not in tail


def test_kernels_under_the_reversed_schedule():
    """The emulation runs one warp at a time, lowest first; CUSIM_REVERSE visits warps and lanes highest first.  A missing barrier
    between a producer and a consumer warp gives stale data under one of the two orders."""
    _run({"CUSIM_REVERSE": "1"}, "-k", "mesh_matches_oracle or light_stage or set_blocks_on_device or device_sink")


def test_upload_paths_under_a_busy_copy_engine():
    """CUSIM_EVENT_BUSY makes cudaEventQuery answer 'not ready' a few times first, so the uploading thread of a threaded upload waits
    for the scan threads instead of taking every work item itself: mixed scanned / unscanned blocks."""
    _run({"CUSIM_EVENT_BUSY": "40"}, "-k", "pipelined_world_load or hinted_upload")
