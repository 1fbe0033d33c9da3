"""TEST INFRASTRUCTURE ONLY: CPU emulation of the CUDA kernels (see hostsim_api.cpp)."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_engine = None


def hostsim_engine():
    """Engine bound to tests/hostsim/libilqc_hostsim.so on CPU tensors (built on demand with g++)."""
    global _engine
    if _engine is None:
        from ilqc_b200 import _lib, build
        from ilqc_b200.engine import Engine
        path = build.build_hostsim(verbose=False)
        _engine = Engine(_lib.bind(path), torch.device("cpu"))
    return _engine
