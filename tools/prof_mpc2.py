"""GPU box: cProfile of run_mpc_batched (host-side overhead per window)."""
import cProfile
import os
import pstats
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ilqc_b200.engine import default_engine  # noqa: E402
from ilqc_b200.envs import make_env  # noqa: E402
from ilqc_b200.model_predictive_control import run_mpc_batched  # noqa: E402
eng = default_engine()
env = make_env(dict(env="real_car", track="complex", vref=2.5, reg_cont=1.0, reg_lag=1.0, reg_speed=1.0))
B = int(os.environ.get("MPC_B", 4096))
x0 = env.init_state.reshape(1, -1).repeat(B, 1).cuda()
run_mpc_batched(env, full_horizon=42, window_size=40, overlap=39, max_iter=10, x0=x0, engine=eng)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
run_mpc_batched(env, full_horizon=69, window_size=40, overlap=39, max_iter=10, x0=x0, engine=eng)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
