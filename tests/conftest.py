"""pytest configuration: `gpu` marker = needs a CUDA device (run with -m gpu on the B200 box)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); the parity tests proper")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_mixed():
    from oracle.oracle import Oracle
    o = Oracle(set_Nc=100.0, iiwarm=False, l_sediment=True)
    yield o
    o.close()


@pytest.fixture(scope="session")
def oracle_warm():
    from oracle.oracle import Oracle
    o = Oracle(set_Nc=50.0, iiwarm=True, l_sediment=True)
    yield o
    o.close()


@pytest.fixture(scope="session")
def gpu_mixed():
    from kid_b200.kidmp import Thompson
    t = Thompson(set_Nc=100.0, iiwarm=False, l_sediment=True)
    yield t
    t.close()


@pytest.fixture(scope="session")
def gpu_warm():
    from kid_b200.kidmp import Thompson
    t = Thompson(set_Nc=50.0, iiwarm=True, l_sediment=True)
    yield t
    t.close()
