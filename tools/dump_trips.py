"""GPU box: per-problem line-search trip counts and step-size states of the bench workload (tuning aid for
the candidate-count predictor).  python tools/dump_trips.py -> gpurun_out/trips.npz"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_batch  # noqa: E402
from ilqc_b200.engine import default_engine  # noqa: E402

B = int(os.environ.get("DUMP_B", 8192))
env, x0, w = make_batch(B, 1236)
eng = default_engine()
r = eng.solve(env, "ddp_linquad_reg", 20, x0=x0, weights=w, record_trips=True)
np.savez_compressed("gpurun_out/trips.npz", trips=r.ls_trips.cpu().numpy(), step=r.stepsize_reported.cpu().numpy(),
                    cost=r.cost.cpu().numpy(), gn=r.norm_grad_obj.cpu().numpy())
print("ok", r.ls_trips.shape)
