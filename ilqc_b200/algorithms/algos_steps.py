"""Oracles, `min_step` and `backtracking_line_search` with the reference's names and signatures
(algorithms/algos_steps.py:22-437), for ONE problem, on top of the CUDA kernels.

The reference builds its oracles as closures over autograd graphs; here every closure is one
`ilqc_backward` + one `ilqc_rollout` call (batch of one, one candidate) on the expansion that
`DiffEnv.forward(cmd, approx)` recorded on `traj[0]` / `costs[0]`, and `obj(cmd)` is one `ilqc_rollout_cost`.
The batched engine (`ilqc_solve`) fuses exactly this loop for B problems and L candidates at once; these
functions exist so that code written against the reference's lower-level API (tests/test_newton.py,
tests/test_steps.py, notebooks) runs with only its import lines changed.
"""
import math
from typing import Callable, Union

import numpy as np
import torch

from .. import _lib
from ..envs.backward import lin_quad_backward, quad_backward_newton, quad_backward_ddp
from ..envs.rollout import roll_out_lin, roll_out_exact

Scalar = Union[float, torch.Tensor]
Oracle = Callable[[Scalar], tuple]


def _info(obj):
    info = getattr(obj[0], "_ilqc", None)
    if info is None:
        raise ValueError("traj / costs must come from ilqc_b200 DiffEnv.forward")
    return info


def gd_oracle(costs, cmd) -> Oracle:
    """algos_steps.py:22-43: oracle(s) = (-s grad, -s/2 |grad|^2) with grad = d sum(costs) / d cmd (adjoint kernel)."""
    info = _info(costs)
    env = info["env"]
    eng = env._engine()
    H = info["cmd"].shape[0]
    ex = eng.linearize(env, info["x0"], torch.as_tensor(cmd, dtype=torch.float64).detach().reshape(1, H, -1), t0=info["t0"])
    grad, _ = eng.grad_obj(ex)
    grad = grad[0].cpu()
    fixed_obj = -0.5 * torch.sum(grad * grad)

    def oracle(stepsize):
        return -stepsize * grad, stepsize * fixed_obj

    return oracle


def newton_oracle(env, cmd, step_mode: str = "dir") -> Oracle:
    """algos_steps.py:46-102: Newton step on the flattened problem with the dense (H nu x H nu) Hessian of the total
    cost, assembled and solved on the device by `ilqc_newton_dense` (csrc/newton_dense.inl) instead of autodiff."""
    eng = env._engine()
    cmd_t = torch.as_tensor(cmd, dtype=torch.float64).detach()
    H, nu = cmd_t.shape
    x0 = torch.as_tensor(env.init_state, dtype=torch.float64).detach().reshape(1, -1)
    t0 = env.init_time_iter

    def solve(reg):
        d, opt = eng.newton_dense(env, x0, cmd_t.reshape(1, H, nu), torch.tensor([float(reg)], dtype=torch.float64), t0=t0)
        return d[0].cpu(), opt[0].cpu()

    if "dir" in step_mode:
        reg_ctrl = 0.0
        cmd_newton, newton_opt = solve(reg_ctrl)
        feasible = bool(newton_opt < 0.0)
        while not feasible:
            reg_ctrl = 10.0 * reg_ctrl if reg_ctrl > 0.0 else 1e-6
            cmd_newton, newton_opt = solve(reg_ctrl)
            feasible = bool(newton_opt < 0.0)
            if reg_ctrl > 1e300:       # (the reference loops forever on NaN; stop instead)
                break
        if reg_ctrl > 1e-6:
            print(f"Newton needed to regularize by reg_ctrl={reg_ctrl}")
            if reg_ctrl > 1e6:
                print("Failed to find a descent direction for newton")

        def oracle(stepsize):
            return stepsize * cmd_newton, stepsize * newton_opt

    elif "reg" in step_mode:

        def oracle(stepsize):
            # the reference solves (H + reg I) d = +grad with an LDL factorisation and returns that d and
            # 1/2 d.grad unchanged (algos_steps.py:88-95): sign and all are kept
            d, opt = solve(1 / stepsize)
            return -d, -opt

    else:
        raise NotImplementedError(f"step_mode {step_mode} not implemented yet")
    return oracle


def _escalated_backward(backward, traj, costs, approx, step_mode, handle_bad_dir, label):
    reg_ctrl = 0.0
    gains, opt_step, feasible = backward(traj, costs, reg_ctrl, handle_bad_dir=handle_bad_dir)
    while not feasible and reg_ctrl < 1e6:
        reg_ctrl = 10.0 * reg_ctrl if reg_ctrl > 0.0 else 1e-6
        gains, opt_step, feasible = backward(traj, costs, reg_ctrl, handle_bad_dir=handle_bad_dir)
    if reg_ctrl > 1e-6:
        print(f"{label} {approx} {step_mode} needed to regularize by reg_ctrl={reg_ctrl}")
        if reg_ctrl > 1e6:
            print("Failed to find a descent direction")
    return gains, opt_step, feasible


def classic_oracle(traj, costs, approx: str = "linquad", step_mode: str = "reg", handle_bad_dir: str = "flag") -> Oracle:
    """algos_steps.py:105-178: Gauss-Newton ('linquad') / Newton ('quad') oracles through the Riccati kernels."""
    if approx == "linquad":
        backward = lin_quad_backward
    elif approx == "quad":
        backward = quad_backward_newton
    else:
        raise NotImplementedError
    if step_mode in ["dir", "dirvar"]:
        gains, opt_step, feasible = _escalated_backward(backward, traj, costs, approx, step_mode, handle_bad_dir, "classic")
        diff_cmd = roll_out_lin(traj, gains) if feasible else None

        def oracle(stepsize):
            return (stepsize * diff_cmd, stepsize * opt_step) if feasible else (None, None)

    elif step_mode in ["reg", "regvar"]:

        def oracle(stepsize):
            reg = 1 / stepsize
            gains, opt_step, feasible = backward(traj, costs, reg, handle_bad_dir=handle_bad_dir)
            return (roll_out_lin(traj, gains), opt_step) if feasible else (None, None)

    else:
        raise NotImplementedError
    return oracle


def ddp_oracle(env, traj, costs, cmd, approx: str = "linquad", step_mode: str = "reg", handle_bad_dir: str = "flag") -> Oracle:
    """algos_steps.py:181-274: DDP oracles (exact roll-out under the gains of the Riccati recursion)."""
    if approx == "linquad":
        backward = lin_quad_backward
    elif approx == "quad":
        backward = quad_backward_ddp
    else:
        raise NotImplementedError
    if step_mode in ["dir", "dirvar"]:
        gains, opt_step, feasible = _escalated_backward(backward, traj, costs, approx, step_mode, handle_bad_dir, "ddp")

        def oracle(stepsize):
            diff_cmd = roll_out_exact(env, traj, gains, cmd, stepsize=stepsize) if feasible else None
            dec_obj = stepsize * opt_step if feasible else None
            return diff_cmd, dec_obj

    elif step_mode in ["reg", "regvar"]:

        def oracle(stepsize):
            reg = 1 / stepsize
            gains, opt_step, feasible = backward(traj, costs, reg, handle_bad_dir=handle_bad_dir)
            return (roll_out_exact(env, traj, gains, cmd, stepsize=1.0), opt_step) if feasible else (None, None)

    elif step_mode == "regstate":

        def oracle(stepsize):
            gains, opt_step, feasible = backward(traj, costs, reg_state=1 / stepsize, handle_bad_dir=handle_bad_dir)
            return (roll_out_exact(env, traj, gains, cmd, stepsize=1.0), opt_step) if feasible else (None, None)

    else:
        raise NotImplementedError
    return oracle


def min_step(env, traj, costs, cmd, stepsize, line_search: bool, algo: str, handle_bad_dir: str):
    """algos_steps.py:277-381: one step of `algo` from (cmd, traj, costs, stepsize)."""
    cmd = torch.as_tensor(cmd, dtype=torch.float64).detach()
    if algo == "gd":
        oracle = gd_oracle(costs, cmd)
        approx = None
        step_mode = "regvar"
    elif "newton" in algo:
        algo_type, step_mode = algo.split("_")
        approx = None
        oracle = newton_oracle(env, cmd)
    else:
        algo_type, approx, step_mode = algo.split("_")
        if algo_type == "classic":
            oracle = classic_oracle(traj, costs, approx, step_mode, handle_bad_dir)
        elif algo_type == "ddp":
            oracle = ddp_oracle(env, traj, costs, cmd, approx, step_mode, handle_bad_dir)
        else:
            raise NotImplementedError

    def obj(c):
        _, cs = env.forward(c)
        return sum(cs)

    if line_search:
        if step_mode == "dir":
            stepsize, decrease_fac, slope_rtol = 1.0, 0.9, 1.0
        elif step_mode == "reg":
            increase_fac = 8
            norm_grad_obj = torch.sqrt(sum([c.grad_ctrl.dot(c.grad_ctrl) + c.grad_state.dot(c.grad_state) for c in costs[1:]]))
            stepsize = increase_fac * stepsize / norm_grad_obj
            decrease_fac, slope_rtol = 0.5, 1.0
        elif step_mode == "regvar":
            stepsize *= 8
            decrease_fac, slope_rtol = 0.5, 1.0
        elif step_mode == "dirvar":
            stepsize = min(4 * stepsize, 1)
            decrease_fac, slope_rtol = 0.9, 1.0
        else:
            raise NotImplementedError
        diff_cmd, stepsize = backtracking_line_search(obj, cmd, oracle, stepsize, decrease_fac, slope_rtol)
        if step_mode == "reg":
            stepsize = stepsize * norm_grad_obj
    else:
        diff_cmd, _ = oracle(stepsize)
    if diff_cmd is not None:
        next_cmd = cmd + diff_cmd
        traj, costs = env.forward(next_cmd, approx=approx)
    else:
        next_cmd = traj = costs = None
    return next_cmd, traj, costs, stepsize


def backtracking_line_search(fun, var, oracle, stepsize, decrease_fac, slope_rtol=1.0):
    """algos_steps.py:384-437, statement for statement (strict `<`, NaN rejects, stop thresholds 1e-32)."""
    dir = float(np.sign(float(stepsize)))
    tol = 0.0

    curr_val = fun(var)
    diff_var, dec_obj = oracle(stepsize)

    if diff_var is None:
        decrease = False
        stop = abs(float(stepsize)) < 1e-32
    else:
        new_val = fun(var + diff_var)
        decrease = bool(dir * (new_val - (curr_val + slope_rtol * dec_obj)) < tol)
        stop = bool(torch.norm(diff_var) < 1e-32)

    while not decrease and not stop:
        stepsize *= decrease_fac
        diff_var, dec_obj = oracle(stepsize)
        if diff_var is None:
            decrease = False
            stop = abs(float(stepsize)) < 1e-32
        else:
            new_val = fun(var + diff_var)
            decrease = bool(dir * (new_val - (curr_val + slope_rtol * dec_obj)) < tol)
            stop = bool(torch.norm(diff_var) < 1e-32)

    if stop:
        print("line-search failed")
    return diff_var, stepsize
