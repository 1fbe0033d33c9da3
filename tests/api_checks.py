"""Checks of the reference-named Python API (run_min_algo, env.forward(cmd, approx), lin_quad_backward, roll_out_*,
mpc_step, run_mpc_batched).  The API always goes through `ilqc_b200.engine.default_engine()`: the CUDA library on the
GPU box (tests/test_gpu_api.py), the test-only host simulator when tests/test_host_logic.py installs it as the default
engine.  Golden vectors come from the unmodified reference (tests/golden/make_golden*.py)."""
import numpy as np
import torch

from tests import parity_checks as pc


def check_run_min_algo_dropin_and_restart():
    """README-style call: run_min_algo(env, algo=..., max_iter=...) -> (cmd_opt, aux, metrics); restart with
    prev_cmd / optim_aux_vars / past_metrics continues the same trace (run_min_algo.py:60-75)."""
    from ilqc_b200.envs import Pendulum
    from ilqc_b200.algorithms import run_min_algo
    g = pc.golden("run_pend_h100_ddp_linquad_reg")
    env = Pendulum(horizon=100, dt=0.02)
    cmd, aux, m = run_min_algo(env, max_iter=8, algo="ddp_linquad_reg")
    assert pc.relerr(torch.tensor(m["cost"]), g["cost"][:9]) < 1e-9 and m["iteration"] == list(range(9))
    assert pc.relerr(cmd, g["cmds"][8]) < 1e-9 and abs(aux["stepsize"] - g["stepsize_state"][7]) < 1e-9 * aux["stepsize"]
    cmd2, aux2, m2 = run_min_algo(env, max_iter=12, algo="ddp_linquad_reg", prev_cmd=cmd, optim_aux_vars=aux, past_metrics=m)
    assert m2["iteration"] == list(range(13))
    assert pc.relerr(torch.tensor(m2["cost"]), g["cost"][:13]) < 1e-9
    assert pc.relerr(torch.tensor(m2["stepsize"]), g["stepsize_reported"][:13]) < 1e-7
    assert pc.relerr(cmd2, g["cmds"][12]) < 1e-9
    try:
        run_min_algo(env, algo="ddp_cubic_reg")
    except AssertionError:
        pass
    else:
        raise AssertionError("bad algo string accepted")


def check_reference_style_lowlevel_api():
    """tests/test_steps.py of the reference, line for line, on the engine: the four direction norms."""
    from ilqc_b200.envs import Car
    from ilqc_b200.envs.backward import lin_quad_backward, quad_backward
    from ilqc_b200.envs.rollout import roll_out_lin, roll_out_exact
    g = pc.golden("steps_car")
    env = Car(model="simple", discretization="euler", cost="exact", horizon=50, dt=0.02)
    ctrls = torch.tensor(g["ctrls"])
    traj, costs = env.forward(ctrls, approx="linquad")
    policies = lin_quad_backward(traj, costs)[0]
    gn = roll_out_lin(traj, policies)
    ddp_lq = roll_out_exact(env, traj, policies, ctrls)
    traj, costs = env.forward(ctrls, approx="quad")
    policies = quad_backward(traj, costs, reg_ctrl=10.0)[0]
    nt = roll_out_lin(traj, policies)
    ddp_q = roll_out_exact(env, traj, policies, ctrls)
    for mine, key in ((gn, "gauss_newton_dir"), (ddp_lq, "ddp_linquad_dir"), (nt, "newton_dir"), (ddp_q, "ddp_quad_dir")):
        assert pc.relerr(mine, g[key]) < 1e-9
    assert np.allclose([float(torch.norm(v)) for v in (gn, ddp_lq, nt, ddp_q)], g["norms"], rtol=1e-9)
    # expansions as attributes, reference conventions (transposed Jacobians, costs[0] zero)
    gl = pc.golden("lin_scar_euler_exact_h50")
    env = pc.env_of(gl)
    env.init_state = torch.tensor(gl["x0"])
    traj, costs = env.forward(torch.tensor(gl["cmd"]), approx="quad")
    assert pc.relerr(traj[3].grad_dyn_state, gl["grad_dyn_state"][2]) < 1e-10
    assert pc.relerr(costs[3].hess_state, gl["hess_state"][3]) < 1e-10 and float(costs[0]) == 0.0
    assert pc.relerr(traj[3].hess_dyn_state(torch.tensor(gl["lam"][0, 2])), gl["hess_dyn_state"][0, 2]) < 1e-10



def check_window(metrics, cmd_opt, cost_ref, cmd_ref, tag):
    """Per-iteration costs of one MPC window.  Strict (1e-9) on the first iterations; near convergence the
    regularisation 1/s becomes tiny, one iteration amplifies 1e-14 input noise to ~1e-9 and a flipped
    line-search trip changes the rest of the free-running window, so the tail of the
    10-iteration window is checked loosely."""
    c = np.array(metrics["cost"])
    # (a window stops early when two consecutive costs are bit-identical, run_min_algo.py:248-256: that
    # test is rounding noise, so the number of recorded iterations may differ by a couple at the tail)
    assert abs(len(c) - len(cost_ref)) <= 3 and min(len(c), len(cost_ref)) >= 6, (tag, len(c), len(cost_ref))
    n = min(len(c), len(cost_ref))
    c, cost_ref = c[:n], cost_ref[:n]
    assert np.max(np.abs(c[:6] - cost_ref[:6]) / np.abs(cost_ref[:6])) < 1e-9, tag
    assert np.max(np.abs(c - cost_ref) / np.abs(cost_ref)) < 1e-3, tag
    assert cmd_opt.shape == cmd_ref.shape, tag
    assert pc.relerr(cmd_opt, cmd_ref) < 5e-2, tag


def check_mpc_step_dropin_against_live_reference_windows(engine):
    """mpc_step chained like run_mpc on the complex track: windows 0..3 of the live reference run
    (tests/golden/mpc_complex_live.npz), each warm-started from the reference's previous window (teacher
    forcing), including the negative-index quirk of window 1 (previous horizon 20 < overlap 39)."""
    from ilqc_b200.envs import make_env
    from ilqc_b200.model_predictive_control import mpc_step, run_mpc_batched
    g = pc.golden("mpc_complex_live")
    cfg = pc.cfg_of(g)
    for w in range(4):
        env = make_env(cfg)
        prev = None if w == 0 else torch.tensor(g[f"cmd_opt_{w - 1}"])
        cmd_opt, metrics = mpc_step(env, full_horizon=int(g[f"full_horizon_{w}"]), overlap=39, max_iter=10,
                                    algo="ddp_linquad_reg", prev_cmd=prev)
        check_window(metrics, cmd_opt, g[f"cost_{w}"], g[f"cmd_opt_{w}"], w)
        assert pc.relerr(env.init_state, g[f"x0_{w}"]) < 1e-9 and env.init_time_iter == int(g[f"t0_{w}"])
    # the batched closed loop walks the same horizons and re-anchors the same way (3 identical episodes)
    env = make_env(cfg)
    r = run_mpc_batched(env, full_horizon=43, window_size=40, overlap=39, max_iter=10, batch=3, engine=engine,
                        expansion=0)   # (same expansion kernel as the drop-in mpc_step: the two modes differ by FMA rounding on the device)
    assert r.horizons == [40, 41, 42, 43] and r.cmd.shape == (3, 43, 3)
    assert torch.equal(r.cmd[0], r.cmd[2])
    assert abs(float(r.cost[0, 0, 0]) - g["cost_0"][0]) < 1e-9 * g["cost_0"][0]
    assert pc.relerr(r.cost[0, 0, :6], g["cost_0"][:6]) < 1e-9
    # ... and is bit-identical to the drop-in mpc_step chained through the host (the free-running chain
    # itself drifts away from the reference's after ~10 near-converged iterations, see _check_window)
    prev = None
    for h in r.horizons:
        prev, _ = mpc_step(make_env(cfg), full_horizon=h, overlap=39, max_iter=10, prev_cmd=prev)
    assert torch.equal(prev, r.cmd[1].cpu())


MPC_VARIANTS = [dict(optim_on_full_window=True), dict(keep_applying_last_ctrl=False), dict(window_size=12, overlap=8)]


def check_mpc_batched_variants_equal_chained_mpc_step(engine, variant):
    """run_mpc_batched with optim_on_full_window / keep_applying_last_ctrl=False / a sliding time > 1
    (mpc_pipeline.py:24-41, :101-106) is bit-identical to the drop-in mpc_step chained through the host."""
    from ilqc_b200.envs import make_env
    from ilqc_b200.model_predictive_control import mpc_step, run_mpc_batched
    cfg = dict(env="simple_car", track="simple", horizon=12, dt=0.04)
    v = dict(variant)
    window, overlap = v.pop("window_size", 12), v.pop("overlap", 11)
    r = run_mpc_batched(make_env(cfg), full_horizon=window + 2 * (window - overlap), window_size=window, overlap=overlap,
                        max_iter=3, batch=2, engine=engine, expansion=0, **v)
    prev = None
    for h in r.horizons:
        prev, _ = mpc_step(make_env(cfg), full_horizon=h, overlap=overlap, max_iter=3, prev_cmd=prev, **v)
    assert len(r.horizons) == 3 and r.cmd.shape[1] == r.horizons[-1]
    assert torch.equal(prev, r.cmd[1].cpu())


def check_mpc_cached_windows_teacher_forced():
    """Cached complex-track MPC files of the reference (valid at HEAD): window h -> h+1 from the cached cmd_opt."""
    from ilqc_b200.envs import make_env
    from ilqc_b200.model_predictive_control import mpc_step
    g = pc.golden("cached_mpc_complex")
    cfg = dict(env="real_car", track="complex", vref=2.5, reg_cont=1.0, reg_lag=1.0, reg_speed=1.0)
    for h in (42, 43, 100):
        env = make_env(cfg)
        cmd_opt, metrics = mpc_step(env, full_horizon=h + 1, overlap=39, max_iter=10, prev_cmd=torch.tensor(g[f"cmd_opt_{h}"]))
        check_window(metrics, cmd_opt, g[f"cost_{h + 1}"], g[f"cmd_opt_{h + 1}"], h)
        # the frozen prefix is spliced unchanged (mpc_pipeline.py:53)
        assert torch.equal(cmd_opt[:h - 39], torch.tensor(g[f"cmd_opt_{h}"])[:h - 39])
