import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_py as O
    O.build()
    return O


@pytest.fixture(scope="session")
def hostsim(oracle):
    """Test-only CPU execution of the kernel cores (tests/hostsim); never used by the product."""
    import _parity
    return _parity.hostsim_lib()


@pytest.fixture(scope="session")
def cuda_lib():
    """The product library; fails loudly when it is missing or no GPU is usable."""
    from dlsc_gc_planner_b200 import capi
    lib = capi.load_library()
    assert lib.dlsc_device_count() > 0, "no CUDA device visible"
    return lib
