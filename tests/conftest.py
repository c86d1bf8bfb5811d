import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


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
