"""solve_ctrl_pb(env_cfg, optim_cfg, ...): make_env + run_min_algo, the call signature of
finite_horizon_control/fhc_pipeline.py:16-35.  (The on-disk experiment cache of the reference,
utils_pipeline/, is host-side bookkeeping outside the accelerated path.)"""
from typing import Any, Optional

import torch

from ..algorithms.run_min_algo import run_min_algo
from ..envs import make_env


def solve_ctrl_pb(env_cfg: dict[str, Any], optim_cfg: dict[str, Any], prev_cmd: Optional[torch.Tensor] = None,
                  optim_aux_vars: Optional[dict[str, Any]] = None, past_metrics: Optional[dict[str, Any]] = None):
    env = make_env(env_cfg)
    cmd_opt, optim_aux_vars, metrics = run_min_algo(env, **optim_cfg, prev_cmd=prev_cmd,
                                                    optim_aux_vars=optim_aux_vars, past_metrics=past_metrics)
    return dict(cmd_opt=cmd_opt, optim_aux_vars=optim_aux_vars, metrics=metrics)
