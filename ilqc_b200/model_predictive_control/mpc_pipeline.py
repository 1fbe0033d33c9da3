"""Sliding-window (nominal) MPC on top of the CUDA engine.

`mpc_step` is the drop-in of model_predictive_control/mpc_pipeline.py:18-55 for one episode (same
arguments, same warm start: last `overlap` controls + copies of the last control, re-anchoring of
`env.init_state` / `env.init_time_iter`, same splice of the result, including Python's negative-index
behaviour when the previous horizon is shorter than the overlap).

`run_mpc_batched` runs B closed-loop episodes concurrently: the loop of run_mpc (mpc_pipeline.py:87-115)
with the state carried on the device instead of through the reference's on-disk experiment cache.  The
plant is the model (no noise), as in the reference.
"""
from dataclasses import dataclass
from typing import Any, Optional

import torch

from ..algorithms.run_min_algo import run_min_algo, run_min_algo_batched
from ..engine import default_engine
from ..envs import make_env


def mpc_step(env, full_horizon: int, overlap: int, max_iter: int = 10, algo: str = "ddp_linquad_reg",
             optim_on_full_window: bool = False, keep_applying_last_ctrl: bool = True,
             prev_cmd: Optional[torch.Tensor] = None):
    curr_horizon = len(prev_cmd) if prev_cmd is not None else 0
    additional_time_steps = full_horizon - curr_horizon
    if prev_cmd is not None:
        prev_cmd = torch.as_tensor(prev_cmd, dtype=torch.float64).detach()
        if keep_applying_last_ctrl:
            additional_cmd = prev_cmd[-1].repeat(additional_time_steps, 1)
        else:
            additional_cmd = torch.zeros(additional_time_steps, env.dim_ctrl, dtype=torch.float64)
        if optim_on_full_window:
            init_cmd = torch.cat((prev_cmd, additional_cmd))
        else:
            init_cmd = torch.cat((prev_cmd[curr_horizon - overlap:], additional_cmd))
            traj, _ = env.forward(prev_cmd)
            env.init_state = traj[curr_horizon - overlap].detach().clone()
            env.init_time_iter = curr_horizon - overlap
        cmd, _, metrics = run_min_algo(env, max_iter, algo, prev_cmd=init_cmd)
    else:
        cmd, _, metrics = run_min_algo(env, max_iter, algo)
    if prev_cmd is None or optim_on_full_window:
        cmd_opt = cmd
    else:
        cmd_opt = torch.cat((prev_cmd[:-overlap], cmd))
    return cmd_opt, metrics


def run_mpc_step(env_cfg: dict[str, Any], optim_cfg: dict[str, Any], prev_cmd: Optional[torch.Tensor] = None):
    env = make_env(env_cfg)
    cmd_opt, metrics = mpc_step(env, prev_cmd=prev_cmd, **optim_cfg)
    return dict(cmd_opt=cmd_opt, metrics=metrics)


@dataclass
class MpcResult:
    cmd: torch.Tensor        # [B, full_horizon, nu] applied + planned controls (cmd_opt of the last window)
    cost: torch.Tensor       # [B, n_windows, max_iter+1] cost trace of every window
    status: torch.Tensor     # [B, n_windows] solver status of every window
    horizons: list           # full_horizon of every window
    n_iterations: int        # total solver iterations executed (sum over episodes and windows)


def run_mpc_batched(env, full_horizon: int, window_size: int, overlap: int, max_iter: int = 10,
                    algo: str = "ddp_linquad_reg", x0=None, weights=None, batch=None, n_candidates: int = 128,
                    engine=None, max_windows: Optional[int] = None, optim_on_full_window: bool = False,
                    keep_applying_last_ctrl: bool = True, expansion: Optional[int] = None) -> MpcResult:
    """B concurrent episodes of run_mpc (mpc_pipeline.py:87-115): horizon <- min(max(horizon + window_size -
    overlap, window_size), full_horizon) per window, each window = mpc_step with the previous window's result.
    The first window is solved on `env.horizon` steps like the reference (run_min_algo sizes cmd from
    env.horizon, run_min_algo.py:53).  `optim_on_full_window` re-optimises all `horizon` controls of every window
    from the initial state; `keep_applying_last_ctrl=False` pads the warm start with zeros (mpc_pipeline.py:24-41)."""
    eng = engine or default_engine()
    dev = eng.device
    if x0 is None:
        B = batch or 1
        x0 = env.init_state.reshape(1, -1).repeat(B, 1)
    x0 = torch.as_tensor(x0, dtype=torch.float64).to(dev)
    B = x0.shape[0]
    wts = None if weights is None else torch.as_tensor(weights, dtype=torch.float64).to(dev)
    sliding = window_size - overlap
    if expansion is None:
        # windows of a few thousand episodes do not fill the GPU: expand every time step of every episode at once
        # (ilqc_algo_desc.expansion = 1) instead of walking each window's horizon in one thread
        expansion = 1 if B <= 8192 else 0
    horizon, costs, stats, horizons, n_it = 0, [], [], [], 0
    cmd_hist = None                      # [B, T, nu]
    k_frozen, x_frozen = 0, x0           # state after the first k_frozen controls of cmd_hist
    while horizon < full_horizon:
        horizon = min(max(horizon + sliding, window_size), full_horizon)
        if cmd_hist is None:
            r = run_min_algo_batched(env, max_iter, algo, x0=x0, weights=wts, t0=0, n_candidates=n_candidates, engine=eng,
                                     expansion=expansion)
            cmd_hist = r.cmd
        else:
            curr = cmd_hist.shape[1]
            extra = horizon - curr
            pad = cmd_hist[:, -1:].repeat(1, extra, 1) if keep_applying_last_ctrl else cmd_hist.new_zeros(B, extra, cmd_hist.shape[2])
            if optim_on_full_window:
                # the env keeps its initial state and time (mpc_pipeline.py:35-36); the window is the whole horizon
                r = run_min_algo_batched(env, max_iter, algo, x0=x0, prev_cmd=torch.cat((cmd_hist, pad), dim=1), weights=wts,
                                         t0=0, n_candidates=n_candidates, engine=eng, expansion=expansion)
                cmd_hist = r.cmd
            else:
                init_cmd = torch.cat((cmd_hist[:, curr - overlap:], pad), dim=1)
                a = curr - overlap       # anchor index into the trajectory of cmd_hist (may be negative: Python slicing)
                a_pos = a if a >= 0 else curr + 1 + a
                if a_pos < k_frozen:
                    k_frozen, x_frozen = 0, x0
                if a_pos > k_frozen:     # advance the frozen prefix with the plant = model
                    traj, _, _ = eng.rollout_cost(env, x_frozen, cmd_hist[:, k_frozen:a_pos], t0=k_frozen, weights=wts)
                    x_frozen, k_frozen = traj[:, -1].contiguous(), a_pos
                r = run_min_algo_batched(env, max_iter, algo, x0=x_frozen, prev_cmd=init_cmd, weights=wts, t0=a,
                                         n_candidates=n_candidates, engine=eng, expansion=expansion)
                cmd_hist = torch.cat((cmd_hist[:, :-overlap], r.cmd), dim=1)  # prev_cmd[:-overlap] (mpc_pipeline.py:53)
        costs.append(r.cost)
        stats.append(r.status)
        horizons.append(horizon)
        n_it += int(r.n_done.sum().item())
        if max_windows is not None and len(horizons) >= max_windows:
            break
    return MpcResult(cmd=cmd_hist, cost=torch.stack(costs, dim=1), status=torch.stack(stats, dim=1),
                     horizons=horizons, n_iterations=n_it)
