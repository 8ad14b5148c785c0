import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    return oracle_lib.load()


@pytest.fixture(scope="session")
def cuda_engine_factory():
    """Engines for the -m gpu tests; fails loudly (no fallback) if CUDA or the library is missing."""
    import torch
    import slr_b200
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    engines = []

    def make(W, H, max_batch=1):
        e = slr_b200.Engine(W, H, max_batch=max_batch, device=0)
        engines.append(e)
        return e

    yield make
    for e in engines:
        e.close()
