import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return load


@pytest.fixture(scope="session")
def engine():
    """One Engine for the whole GPU session; fails loudly without a B200 / the library."""
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from datmo_using_optical_flow_b200.engine import Engine
    eng = Engine(0)
    yield eng
    eng.close()
