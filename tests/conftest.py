import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """Build libherbgpu.so (nvcc, sm_100a; cross-compiles without a GPU) if it is missing or older than its sources,
    so the suite also runs from a fresh checkout.  A GPU box without nvcc uses the prebuilt library it was sent."""
    from herbolution_b200 import build as b
    try:
        b.build_cuda()
    except RuntimeError:
        if not os.path.exists(b.LIB):
            raise


@pytest.fixture(scope="session")
def ctx():
    """One libherbgpu context on cuda:0.  GPU tests call the product ONLY through the C ABI."""
    import herbolution_b200 as hb
    c = hb.Context(device=0, seed=0)
    yield c
    c.close()
