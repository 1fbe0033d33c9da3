import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def cuda_engine():
    """The product engine (libilqc_b200.so on cuda:0).  Fails loudly if the CUDA library is missing."""
    from ilqc_b200.engine import default_engine
    assert torch.cuda.is_available(), "gpu-marked test without a CUDA device"
    return default_engine()


@pytest.fixture(scope="session")
def sim_engine():
    """Test-only CPU emulation of the same kernel templates (tests/hostsim)."""
    from tests.hostsim import hostsim_engine
    return hostsim_engine()
