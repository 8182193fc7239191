import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def rt_factory():
    """GPU tests call the product only through the C ABI; if the library or GPU is missing they fail loudly."""
    from luz_b200 import rt as rtmod
    made = []

    def make(**kw):
        r = rtmod.LuzRT(**kw)
        made.append(r)
        return r

    yield make
    for r in made:
        r.close()
