"""Drop-in `run_min_algo` on top of the CUDA engine (algorithms/run_min_algo.py:15-100 of the reference).

`run_min_algo(env, max_iter, algo, ...)` keeps the reference's signature, return types and restart
arguments (`prev_cmd`, `optim_aux_vars`, `past_metrics`) for ONE problem; `run_min_algo_batched` is the
same call for B problems of one family (per-problem initial states / cost weights / time offsets), which is
what the engine is built for.  Both go through `ilqc_solve` in libilqc_b200.so; there is no CPU path.
"""
import math
from typing import Any, Optional

import torch

from .. import _lib
from ..engine import default_engine, default_stepsize


def check_nomenclature(algo: str):
    """Same grammar and assertions as run_min_algo.py:267-292."""
    if algo == "gd":
        return "classic", None, "dir"
    if "newton" in algo:
        algo_type, step_mode = algo.split("_")
        assert step_mode in ["dir", "reg", "dirvar", "regvar"]
        return algo_type, None, step_mode
    algo_type, approx, step_mode = algo.split("_")
    assert algo_type in ["classic", "ddp"]
    assert approx in ["linquad", "quad"]
    assert step_mode in ["dir", "reg", "dirvar", "regvar"]
    return algo_type, approx, step_mode


def check_cvg_status(metrics: dict, verbose: bool = True) -> str:
    """run_min_algo.py:195-264 on a metrics dict of python lists."""
    status = "running"
    c = metrics["cost"]
    if math.isnan(c[-1]) or c[-1] > c[0] + 1e3:
        status = "diverged"
    if len(c) > 1 and abs(c[-1] - c[-2]) / abs(c[-2]) < 1e-24:
        status = "converged"
    if status == "running" and metrics["stepsize"][-1] < 1e-24:
        status = "got stuck"
    if verbose and status != "running":
        print("{0} {1}".format(metrics["algo"][0], status))
    return status


def run_min_algo_batched(env, max_iter: int = 10, algo: str = "ddp_linquad_reg", stepsize=None,
                         line_search: bool = True, handle_bad_dir: str = "check_total_value", x0=None,
                         prev_cmd=None, weights=None, t0=None, batch=None, n_candidates: int = 128,
                         compute_grad_norm: bool = True, engine=None, expansion: int = 0, cost0=None):
    """B problems at once.  x0 [B,nx], prev_cmd [B,H,nu] (None: zeros), stepsize float or [B], weights [B,3]
    (Car reg_cont/reg_lag/reg_speed), t0 int or [B] (init_time_iter).  Returns ilqc_b200.engine.SolveResult
    with device tensors: cmd [B,H,nu], stepsize [B], cost / norm_grad_obj / stepsize_reported [B,max_iter+1],
    n_done [B], status [B]."""
    check_nomenclature(algo)
    if handle_bad_dir not in _lib.BAD_DIR:   # 'modify_hess' needs the removed torch.eig in the reference
        raise NotImplementedError("handle_bad_dir=%r" % handle_bad_dir)
    eng = engine or default_engine()
    return eng.solve(env, algo, max_iter, x0=x0, cmd0=prev_cmd, stepsize=stepsize, batch=batch, weights=weights,
                     t0=env.init_time_iter if t0 is None else t0, n_candidates=n_candidates,
                     line_search=line_search, compute_grad_norm=compute_grad_norm, handle_bad_dir=handle_bad_dir,
                     expansion=expansion, cost0=cost0)


def run_min_algo(env, max_iter: int = 10, algo: str = "ddp_linquad_reg", stepsize: Optional[float] = None,
                 line_search: bool = True, handle_bad_dir: str = "check_total_value", compute_min_sev: bool = False,
                 prev_cmd: Optional[torch.Tensor] = None, optim_aux_vars: Optional[dict[str, Any]] = None,
                 past_metrics: Optional[dict[str, Any]] = None, verbose: bool = False):
    """One problem, reference return convention: (cmd_opt [H,nu] CPU tensor, {'stepsize': float},
    metrics dict of lists with keys iteration/time/cost/norm_grad_obj/stepsize/algo).

    Restart semantics of the reference are kept: with `past_metrics` the solve continues from iteration
    past_metrics['iteration'][-1] up to `max_iter` and appends to the same lists."""
    if compute_min_sev:
        # (the reference itself raises here: `format_types.append["scientific"]`, run_min_algo.py:184;
        #  compute_min_sing_eigval_jac below is the function it would have called)
        raise NotImplementedError("compute_min_sev raises in the reference too (run_min_algo.py:184)")
    import time
    _, _, step_mode = check_nomenclature(algo)
    if "newton" in algo:
        return _run_min_algo_newton(env, max_iter, algo, stepsize, line_search, handle_bad_dir, prev_cmd, optim_aux_vars,
                                    past_metrics)
    if optim_aux_vars is not None:
        stepsize = float(optim_aux_vars["stepsize"])
    elif stepsize is None:
        stepsize = default_stepsize(algo)
    it0 = 0 if past_metrics is None else past_metrics["iteration"][-1]
    if past_metrics is not None and check_cvg_status(past_metrics, verbose=False) != "running":
        cmd = prev_cmd if prev_cmd is not None else torch.zeros(env.horizon, env.dim_ctrl, dtype=torch.float64)
        return cmd.detach(), dict(stepsize=stepsize), past_metrics
    n = max(max_iter - it0, 0)
    cmd0 = None if prev_cmd is None else torch.as_tensor(prev_cmd, dtype=torch.float64).detach().unsqueeze(0)
    tic = time.time()
    # restart: divergence is tested against the ORIGINAL first cost (run_min_algo.py:240-243)
    cost0 = None if past_metrics is None else torch.tensor([past_metrics["cost"][0]], dtype=torch.float64)
    r = run_min_algo_batched(env, n, algo, stepsize, line_search, handle_bad_dir, x0=env.init_state.reshape(1, -1),
                             prev_cmd=cmd0, batch=1, cost0=cost0)
    wall = time.time() - tic
    done = int(r.n_done[0])
    cost, gn, srep = r.cost[0].cpu().tolist(), r.norm_grad_obj[0].cpu().tolist(), r.stepsize_reported[0].cpu().tolist()
    rows = range(0 if past_metrics is None else 1, done + 1)
    if past_metrics is None:
        metrics = dict(iteration=[], time=[], cost=[], norm_grad_obj=[], stepsize=[], algo=[])
        t_prev = 0.0
    else:
        metrics, t_prev = past_metrics, past_metrics["time"][-1]
    for k in rows:
        # (restart: the engine rescales the reported 'reg' step by the gradient norm it recomputes at
        # prev_cmd, which is the last recorded metrics['norm_grad_obj'] of the previous call)
        rep = srep[k]
        metrics["iteration"].append(it0 + k)
        metrics["time"].append(t_prev + wall * k / max(done, 1))
        metrics["cost"].append(cost[k])
        metrics["norm_grad_obj"].append(gn[k])
        metrics["stepsize"].append(rep)
        metrics["algo"].append(algo)
    status = int(r.status[0])
    cmd_opt = r.cmd[0].cpu() if status != _lib.NO_DIRECTION else None
    if verbose:
        for k in rows:
            print("{:<12d}{:<16.6e}{:<16.2e}{:<16.2e}".format(it0 + k, cost[k], gn[k], srep[k]))
    return cmd_opt, dict(stepsize=float(r.stepsize[0])), metrics


def _run_min_algo_newton(env, max_iter, algo, stepsize, line_search, handle_bad_dir, prev_cmd, optim_aux_vars, past_metrics):
    """`newton_*` (run_min_algo.py:50-100 with min_step's newton_oracle branch, algos_steps.py:311-315): the dense Newton
    step on the flattened problem, a diagnostic of the reference for ONE problem.  The loop is the reference's; every
    number comes from the device (`ilqc_newton_dense` for the direction, `ilqc_rollout_cost` for the line search's
    objective, `ilqc_grad_obj` for the metrics)."""
    import time
    from .algos_steps import min_step
    cmd = (torch.as_tensor(prev_cmd, dtype=torch.float64).detach().clone() if prev_cmd is not None
           else torch.zeros(env.horizon, env.dim_ctrl, dtype=torch.float64))
    _, approx, step_mode = check_nomenclature(algo)
    if optim_aux_vars is not None:
        stepsize = optim_aux_vars["stepsize"]
    elif stepsize is None:
        stepsize = 100.0 if "reg" in step_mode else 1.0
    eng = env._engine()

    def info(cmd, costs, iteration, stepsize, t_iter, metrics):
        cost = float(sum(costs)) if costs is not None else float("nan")
        ex = eng.linearize(env, env.init_state.reshape(1, -1), cmd.reshape(1, *cmd.shape), t0=env.init_time_iter)
        gn = float(eng.grad_obj(ex)[1][0])
        row = dict(iteration=iteration, time=(0.0 if iteration == 0 else metrics["time"][-1]) + t_iter, cost=cost,
                   norm_grad_obj=gn, stepsize=stepsize, algo=algo)
        if metrics is None:
            return {k: [v] for k, v in row.items()}
        for k, v in row.items():
            metrics[k].append(v)
        return metrics

    traj, costs = env.forward(cmd, approx=approx)
    if past_metrics is None:
        metrics, iteration = info(cmd, costs, 0, stepsize, 0.0, None), 0
    else:
        metrics, iteration = past_metrics, past_metrics["iteration"][-1]
    status = check_cvg_status(metrics, verbose=False)
    while status == "running" and iteration < max_iter:
        tic = time.time()
        cmd_new, traj, costs, stepsize = min_step(env, traj, costs, cmd, stepsize, line_search, algo, handle_bad_dir)
        iteration += 1
        if cmd_new is None:   # collect_info differentiates a None command in the reference: the run ends here
            metrics = info(cmd, None, iteration, stepsize, time.time() - tic, metrics)
            cmd = None
            break
        cmd = cmd_new
        metrics = info(cmd, costs, iteration, stepsize, time.time() - tic, metrics)
        status = check_cvg_status(metrics, verbose=False)
    return (cmd.detach() if cmd is not None else None), dict(stepsize=stepsize), metrics


def compute_min_sing_eigval_jac(cmd: torch.Tensor, env) -> float:
    """run_min_algo.py:103-124: smallest singular value of G = J J^T, J = d traj[1:] / d cmd (H nx x H nu).  J comes
    from the sensitivity kernel (`ilqc_traj_jacobian`); the Gram product and the singular values are library calls on the
    device.  (G has rank <= H nu < H nx, so the reference's number is rounding noise around zero for every env of the
    toolbox; the Jacobian itself is what the parity test compares.)"""
    eng = env._engine()
    cmd = torch.as_tensor(cmd, dtype=torch.float64).detach()
    jac = eng.traj_jacobian(env, env.init_state.reshape(1, -1), cmd.reshape(1, *cmd.shape), t0=env.init_time_iter)[0]
    s = torch.linalg.svdvals(jac @ jac.t())
    return float(s.min())
