"""GPU box: where does an MPC window go?  4096 episodes, complex track, H=40 windows, 10 iterations."""
import ctypes as C
import os
import sys
import time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import CLASSES  # noqa: E402
from ilqc_b200.engine import default_engine  # noqa: E402
from ilqc_b200.envs import make_env  # noqa: E402
from ilqc_b200.model_predictive_control import run_mpc_batched  # noqa: E402
eng = default_engine()
env = make_env(dict(env="real_car", track="complex", vref=2.5, reg_cont=1.0, reg_lag=1.0, reg_speed=1.0))
B = int(os.environ.get("MPC_B", 4096))
g = torch.Generator().manual_seed(1237)
xi = torch.randn(B, 4, generator=g, dtype=torch.float64)
x0 = env.init_state.reshape(1, -1).repeat(B, 1)
x0[:, 3] = x0[:, 3] + 0.01 * xi[:, 0]
x0[:, 7] = x0[:, 3]
w = (torch.tensor([1.0, 1.0, 1.0], dtype=torch.float64).reshape(1, 3) + 1e-3 * xi[:, 1:]).cuda()
x0 = x0.cuda()
# one window solve alone (H = 40, cold start)
r = eng.solve(env, "ddp_linquad_reg", 10, x0=x0, weights=w, horizon=40)
torch.cuda.synchronize()
ms, n, u = (C.c_double * 8)(), (C.c_int64 * 8)(), (C.c_int64 * 8)()
eng.lib.ilqc_profile_enable(1)
eng.lib.ilqc_profile_read(ms, n, u, 8)
t0 = time.perf_counter()
for _ in range(5):
    r = eng.solve(env, "ddp_linquad_reg", 10, x0=x0, weights=w, horizon=40, cmd0=r.cmd)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
eng.lib.ilqc_profile_read(ms, n, u, 8)
eng.lib.ilqc_profile_enable(0)
print(f"B={B}: warm-started window solve (10 it): {dt * 1e3:.2f} ms wall;", {c: (round(ms[i] / 5, 2), n[i] // 5) for i, c in enumerate(CLASSES)}, flush=True)
t0 = time.perf_counter()
res = run_mpc_batched(env, full_horizon=79, window_size=40, overlap=39, max_iter=10, x0=x0, weights=w, engine=eng)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"run_mpc_batched 40 windows: {dt * 1e3 / 40:.2f} ms per window, {res.n_iterations / dt / 1e6:.2f} M iterations/s", flush=True)
