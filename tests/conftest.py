import os
import subprocess
import sys

import pytest

TESTS = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(TESTS)
for p in (ROOT, os.path.join(ROOT, "oracle"), TESTS):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_ok():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _cuda_ok():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_built():
    """Builds the plain-C oracle (and the reference's own library when its sources are present)."""
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    return True


@pytest.fixture(scope="session")
def native_lib():
    """The product's CUDA library (cross-compiles without a GPU)."""
    from trajectory_generator_b200 import _native
    _native.build_native()
    return _native.lib()


@pytest.fixture(scope="session")
def hostsim(oracle_built):
    import hostsim_loader
    return hostsim_loader.load()
