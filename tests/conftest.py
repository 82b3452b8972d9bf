import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (sm_100a); run with -m gpu")


@pytest.fixture(scope="session")
def cuda():
    """GPU tests call the product through the C ABI.  On a GPU box a missing library or device
    is a FAILURE (never a skip, never a fallback)."""
    import torch
    assert torch.cuda.is_available(), "gpu-marked test running without a CUDA device"
    import spev_tts_b200
    spev_tts_b200.load()
    return torch.device("cuda", 0)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def _load(name):
        return np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return _load
