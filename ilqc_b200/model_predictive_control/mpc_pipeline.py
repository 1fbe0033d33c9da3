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

from ..algorithms.run_min_algo import run_min_algo, run_min_algo_batched, check_cvg_status
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


def check_exp_fail(exp_outputs: dict[str, Any]) -> bool:
    return check_cvg_status(exp_outputs["metrics"], verbose=False) == "diverged"


def run_mpc(env_cfg: dict[str, Any], optim_cfg: dict[str, Any]) -> torch.Tensor:
    """run_mpc of the reference (mpc_pipeline.py:87-115) for one episode: windows of growing `full_horizon`, each
    warm-started from the previous window's cmd_opt, stop at the first diverged window.  The chain is carried in memory
    (the reference stores every window in its on-disk experiment cache, utils_pipeline/).  For many concurrent episodes
    use run_mpc_batched / the C entry ilqc_mpc."""
    from copy import deepcopy
    horizon = 0
    max_horizon, window_size, overlap = optim_cfg["full_horizon"], optim_cfg["window_size"], optim_cfg["overlap"]
    sliding_time = window_size - overlap
    exp_outputs, prev = None, None
    while horizon < max_horizon:
        temp_optim_cfg = deepcopy(optim_cfg)
        horizon = min(max(horizon + sliding_time, window_size), max_horizon)
        del temp_optim_cfg["window_size"]
        temp_optim_cfg["full_horizon"] = horizon
        exp_outputs = run_mpc_step(env_cfg, temp_optim_cfg, prev_cmd=prev)
        prev = exp_outputs["cmd_opt"]
        if check_exp_fail(exp_outputs):
            print("Maximum horizon: {}".format(horizon))
            break
    return exp_outputs["cmd_opt"]


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
    if expansion is None:
        # windows of a few thousand episodes do not fill the GPU: expand every time step of every episode at once
        # (ilqc_algo_desc.expansion = 1) instead of walking each window's horizon in one thread
        expansion = 1 if B <= 8192 else 0
    # the window loop, the warm-start shift and the re-anchoring run inside ilqc_mpc (csrc/mpc.inl)
    cmd, cost, status, n_done = eng.mpc(env, x0, full_horizon, window_size, overlap, max_iter, algo, weights=wts,
                                        n_candidates=n_candidates, optim_on_full_window=optim_on_full_window,
                                        keep_applying_last_ctrl=keep_applying_last_ctrl, max_windows=max_windows or 0,
                                        expansion=expansion)
    horizons, h = [], 0
    sliding = window_size - overlap
    while h < full_horizon and len(horizons) < cost.shape[1]:
        h = min(max(h + sliding, window_size), full_horizon)
        horizons.append(h)
    return MpcResult(cmd=cmd, cost=cost, status=status, horizons=horizons, n_iterations=int(n_done.sum().item()))
