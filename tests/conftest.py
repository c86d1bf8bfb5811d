import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


SIM_DIR = os.path.join(ROOT, "tests", "cusim")
SIM_LIB = os.path.join(SIM_DIR, "_build", "libhcm_sim.so")


def build_sim() -> str:
    """tests/cusim: the CUDA sources of hypercraft_b200/csrc compiled with g++ against a CPU emulation of the CUDA execution
    model (test infrastructure; see tests/cusim/include/cuda_runtime.h)."""
    import subprocess

    subprocess.run(["make", "-C", SIM_DIR], check=True, stdout=subprocess.DEVNULL)
    return SIM_LIB


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # HCM_TEST_SIM=1: run the `gpu` tests against the CPU emulation build of the library instead (developer / CPU-tier use:
    # `HCM_TEST_SIM=1 pytest tests/test_gpu_parity.py -m gpu`).  Never set on the GPU box.
    if os.environ.get("HCM_TEST_SIM"):
        from hypercraft_b200 import mesher

        mesher.LIB_PATH = os.environ.get("HCM_TEST_SIM_LIB") or build_sim()  # HCM_TEST_SIM_LIB: another build of the emulation (sanitizers)
        mesher._lib = None


@pytest.fixture(scope="session")
def registry_blob():
    from hypercraft_b200 import registry

    return registry.load().blob


@pytest.fixture(scope="session")
def port(registry_blob):
    import hcoracle

    hcoracle.build(ref=False, port=True)
    return hcoracle.PortOracle(registry_blob)


@pytest.fixture(scope="session")
def have_ref():
    import hcoracle

    return hcoracle.have_ref()
