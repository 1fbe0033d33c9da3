"""GPU box (or host build with ILQC_SOAK_SIM=1): randomised-configuration parity against the oracle for many seeds
beyond the 16 of the test suite.  python tools/soak_random_configs.py [first_seed] [n_seeds]"""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
torch.set_default_dtype(torch.float64)
from tests import parity_checks as pc  # noqa: E402
if os.environ.get("ILQC_SOAK_SIM"):
    from tests.hostsim import hostsim_engine
    eng = hostsim_engine()
else:
    from ilqc_b200.engine import default_engine
    eng = default_engine()
first = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
bad = 0
worst = 0.0
for seed in range(first, first + n):
    try:
        errs = pc.check_random_config_vs_oracle(eng, seed)
        worst = max(worst, max(errs.values()))
    except Exception as e:  # noqa: BLE001
        bad += 1
        print("FAIL", seed, repr(e)[:500], flush=True)
print(f"seeds {first}..{first + n - 1}: {bad} failures, worst relative error {worst:.2e}", flush=True)
