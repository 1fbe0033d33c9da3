"""solve_ctrl_pb(env_cfg, optim_cfg, ...): make_env + run_min_algo, the call signature of
finite_horizon_control/fhc_pipeline.py:16-35.  `solve_ctrl_pb_incrementally`
(fhc_pipeline.py:46-53) runs the same problem 10 iterations at a time through the restart arguments; the on-disk
experiment cache the reference wraps around it (utils_pipeline/) is host-side bookkeeping outside the accelerated
path, so the chain is carried in memory."""
from typing import Any, Optional

import torch

from ..algorithms.run_min_algo import run_min_algo, check_cvg_status
from ..envs import make_env


def solve_ctrl_pb(env_cfg: dict[str, Any], optim_cfg: dict[str, Any], prev_cmd: Optional[torch.Tensor] = None,
                  optim_aux_vars: Optional[dict[str, Any]] = None, past_metrics: Optional[dict[str, Any]] = None):
    env = make_env(env_cfg)
    cmd_opt, optim_aux_vars, metrics = run_min_algo(env, **optim_cfg, prev_cmd=prev_cmd,
                                                    optim_aux_vars=optim_aux_vars, past_metrics=past_metrics)
    return dict(cmd_opt=cmd_opt, optim_aux_vars=optim_aux_vars, metrics=metrics)


def check_exp_done(exp_outputs: dict[str, Any]) -> bool:
    return check_cvg_status(exp_outputs["metrics"], verbose=False) != "running"


output_to_input = dict(cmd_opt="prev_cmd", optim_aux_vars="optim_aux_vars", metrics="past_metrics")


def solve_ctrl_pb_incrementally(env_cfg: dict[str, Any], optim_cfg: dict[str, Any], param_interval_size: int = 10):
    """run_exp_incrementally_wrapper(solve_ctrl_pb, ..., 'max_iter', 10, ...) of the reference
    (utils_pipeline/record_exp.py:82-130): max_iter <- min(done + 10, max_iter) per leg, every leg restarted from the
    previous leg's (cmd_opt, optim_aux_vars, metrics); a finished experiment (status != 'running') is not relaunched."""
    max_param = optim_cfg["max_iter"]
    param, exp_outputs = 0, None
    while param < max_param:
        leg_cfg = dict(optim_cfg, max_iter=min(param + param_interval_size, max_param))
        if exp_outputs is None:
            exp_outputs = solve_ctrl_pb(env_cfg, leg_cfg)
        elif not check_exp_done(exp_outputs):
            new_inputs = {val: exp_outputs[key] for key, val in output_to_input.items()}
            exp_outputs = solve_ctrl_pb(env_cfg, leg_cfg, **new_inputs)
        param += param_interval_size
    return exp_outputs
