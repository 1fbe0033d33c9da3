"""GPU box: throughput of the other BASELINE.json configurations at their per-GPU sizes (device-resident inputs,
CUDA events, 3 timed repetitions after 1 warm-up).  python tools/bench_configs.py > gpurun_out/bench_configs.txt"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ilqc_b200.engine import default_engine  # noqa: E402
from ilqc_b200.envs import Car, CartPendulum, Pendulum, make_env  # noqa: E402
from ilqc_b200.model_predictive_control import run_mpc_batched  # noqa: E402

eng = default_engine()


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 0
    for _ in range(reps):
        n += fn()
    e1.record()
    torch.cuda.synchronize()
    return n, e0.elapsed_time(e1) * 1e-3


def report(tag, n_it, secs):
    print(f"{tag}: {n_it / secs / 1e6:.2f} M iterations/s ({secs * 1e3 / 3:.1f} ms per solve)", flush=True)


g = torch.Generator().manual_seed(1234)
# config 2: cart-pendulum H=50, B=16384
env = CartPendulum(horizon=50, dt=0.05, stay_put_time=0.6, x_limits=(-2.0, 2.0))
x0 = torch.zeros(16384, 4, dtype=torch.float64)
x0[:, 0] = torch.randn(16384, generator=g, dtype=torch.float64)
x0 = x0.cuda()
for algo in ("classic_linquad_reg", "ddp_linquad_reg"):
    report(f"config 2 cart-pendulum H=50 B=16384 {algo} 20 it", *timed(lambda: int(eng.solve(env, algo, 20, x0=x0).n_done.sum())))
# config 3: simple car exact cost Euler H=50, B=65536, 8 candidates
env = Car(model="simple", track="simple", cost="exact", discretization="euler", reg_bar=0.0, horizon=50, dt=0.04)
x0 = env.init_state.reshape(1, -1).repeat(65536, 1)
x0[:, 3] = 1.0 + 0.01 * torch.randn(65536, generator=g, dtype=torch.float64)
x0[:, 5] = x0[:, 3]
x0 = x0.cuda()
report("config 3 simple car exact Euler H=50 B=65536 ddp_linquad_reg L=8 20 it",
       *timed(lambda: int(eng.solve(env, "ddp_linquad_reg", 20, x0=x0, n_candidates=8).n_done.sum())))
# config 4 nominal algorithm: ddp_quad_reg, B=32768, H=100, 10 iterations
from bench import make_batch  # noqa: E402
env, x0, w = make_batch(32768, 1236)
x0, w = x0.cuda(), w.cuda()
report("config 4 bicycle RK4 H=100 B=32768 ddp_quad_reg 10 it",
       *timed(lambda: int(eng.solve(env, "ddp_quad_reg", 10, x0=x0, weights=w).n_done.sum()), reps=3))
# config 5: MPC, 4096 episodes, complex track, window 40 / overlap 39 / 10 iterations, 80 windows (full_horizon 119)
env = make_env(dict(env="real_car", track="complex", vref=2.5, reg_cont=1.0, reg_lag=1.0, reg_speed=1.0))
x0 = env.init_state.reshape(1, -1).repeat(4096, 1)
xi = torch.randn(4096, 4, generator=g, dtype=torch.float64)
x0[:, 3] = x0[:, 3] + 0.01 * xi[:, 0]
x0[:, 7] = x0[:, 3]
w = torch.tensor([1.0, 1.0, 1.0], dtype=torch.float64).reshape(1, 3) + 1e-3 * xi[:, 1:]
n, secs = timed(lambda: run_mpc_batched(env, full_horizon=119, window_size=40, overlap=39, max_iter=10, x0=x0, weights=w,
                                        engine=eng).n_iterations, reps=1)
print(f"config 5 MPC 4096 episodes x 80 windows x 10 it (H=40): {n / secs / 1e6:.2f} M iterations/s, "
      f"{4096 * 80 / secs:.0f} windows/s, {secs:.2f} s", flush=True)
