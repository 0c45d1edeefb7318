import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    return np.load(os.path.join(ROOT, "tests", "golden", "psiformer_small.npz"))


@pytest.fixture(autouse=True)
def _deterministic_torch_rng():
    """Every test starts from the same torch RNG state (CPU and CUDA), so a test's random inputs do not
    depend on which tests ran before it."""
    import torch

    torch.manual_seed(20261018)
    yield
