"""GPU box: timing of the pieces of one MPC window inside run_mpc_batched (synchronised timers)."""
import os
import sys
import time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ilqc_b200.engine import default_engine  # noqa: E402
from ilqc_b200.envs import make_env  # noqa: E402
from ilqc_b200.algorithms.run_min_algo import run_min_algo_batched  # noqa: E402
eng = default_engine()
env = make_env(dict(env="real_car", track="complex", vref=2.5, reg_cont=1.0, reg_lag=1.0, reg_speed=1.0, horizon=40))
B = 4096
g = torch.Generator().manual_seed(1237)
xi = torch.randn(B, 4, generator=g, dtype=torch.float64)
x0 = env.init_state.reshape(1, -1).repeat(B, 1)
x0[:, 3] = x0[:, 3] + 0.01 * xi[:, 0]
x0[:, 7] = x0[:, 3]
w = (torch.tensor([1.0, 1.0, 1.0], dtype=torch.float64).reshape(1, 3) + 1e-3 * xi[:, 1:]).cuda()
x0 = x0.cuda()
os.makedirs("gpurun_out", exist_ok=True)
T = {}
def tick(name, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    T[name] = T.get(name, 0.0) + time.perf_counter() - t0
    return r
r = run_min_algo_batched(env, 10, "ddp_linquad_reg", x0=x0, weights=w, t0=0, engine=eng)
cmd_hist = r.cmd
k_frozen, x_frozen, horizon = 0, x0, 40
trips = []
for win in range(30):
    horizon += 1
    curr = cmd_hist.shape[1]
    init_cmd = tick("cat_in", lambda: torch.cat((cmd_hist[:, curr - 39:], cmd_hist[:, -1:].repeat(1, horizon - curr, 1)), dim=1))
    a = curr - 39
    traj = tick("prefix_rollout", lambda: eng.rollout_cost(env, x_frozen, cmd_hist[:, k_frozen:a], t0=k_frozen, weights=w)[0])
    x_frozen, k_frozen = traj[:, -1].contiguous(), a
    r = tick("solve", lambda: eng.solve(env, "ddp_linquad_reg", 10, x0=x_frozen, cmd0=init_cmd, weights=w, t0=a, record_trips=True))
    trips.append(r.ls_trips.float().mean().item())
    if win in (3, 12, 25):
        import numpy as np
        np.save(f'gpurun_out/mpc_trips_{win}.npy', r.ls_trips.cpu().numpy())
    cmd_hist = tick("cat_out", lambda: torch.cat((cmd_hist[:, :-39], r.cmd), dim=1))
    tick("item", lambda: int(r.n_done.sum().item()))
print({k: round(v / 30 * 1e3, 2) for k, v in T.items()}, "ms per window; mean trips", sum(trips) / len(trips), "n_done mean", r.n_done.float().mean().item())
