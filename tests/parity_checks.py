"""Parity checks shared by the GPU tests (CUDA engine) and the CPU tests (host simulator of the same
kernel templates).  Every check compares against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py); tolerances are relative, float64.

north_star tolerance: 1e-9 relative on costs and controls per iteration.
"""
import ast
import glob
import os

import numpy as np
import torch

from ilqc_b200 import _lib
from ilqc_b200.envs import make_env

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-9          # north_star: relative, per iteration
TOL_EXPANSION = 1e-10


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def golden_names(prefix):
    return sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def env_of(g):
    return make_env(ast.literal_eval(str(g["env_cfg"])))


def cfg_of(g):
    return ast.literal_eval(str(g["env_cfg"]))


def bad_dir_of(g):
    return str(g["handle_bad_dir"]) if "handle_bad_dir" in g.files else "check_total_value"


def relerr(a, b):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300))


def t64(x):
    return torch.as_tensor(np.asarray(x), dtype=torch.float64)


def is_supported(g):
    return True  # every shipped env / integrator / cost combination of the fixtures is built


# ---- forward / expansions (envs/forward.py) -----------------------------------------------------------
def check_expansion(eng, name):
    g = golden(name)
    env = env_of(g)
    x0, cmd, t0 = t64(g["x0"]).reshape(1, -1), t64(g["cmd"]).unsqueeze(0), int(g["t0"])
    traj, cost_t, total = eng.rollout_cost(env, x0, cmd, t0=t0)
    errs = dict(traj=relerr(traj[0], g["traj"]), costs=relerr(cost_t[0], g["costs"]),
                total=relerr(total[0], g["total_cost"]))
    for i in range(g["lam"].shape[0]):
        ex = eng.expand_dense(env, x0, cmd, t0=t0, lam=t64(g["lam"][i]).unsqueeze(0))
        errs.update({
            "A": relerr(ex["A"][0].transpose(1, 2), g["grad_dyn_state"]),
            "B": relerr(ex["B"][0].transpose(1, 2), g["grad_dyn_ctrl"]),
            "p": relerr(ex["p"][0], g["grad_state"]), "P": relerr(ex["P"][0], g["hess_state"]),
            "q": relerr(ex["q"][0], g["grad_ctrl"]), "Q": relerr(ex["Q"][0], g["hess_ctrl"]),
            f"Hxx{i}": relerr(ex["Hxx"][0][1:], g["hess_dyn_state"][i][1:]),
            f"Hxu{i}": relerr(ex["Hxu"][0], g["hess_dyn_state_ctrl"][i]),
            f"Huu{i}": relerr(ex["Huu"][0], g["hess_dyn_ctrl"][i]),
        })
    bad = {k: v for k, v in errs.items() if not v < TOL_EXPANSION}
    assert not bad, (name, bad)
    return errs


# ---- backward passes and roll-outs (envs/backward.py, envs/rollout.py) --------------------------------
def check_backward_rollout(eng, name):
    g = golden(name)
    env = env_of(g)
    x0, cmd, t0 = t64(g["x0"]).reshape(1, -1), t64(g["cmd"]).unsqueeze(0), int(g["t0"])
    ex = eng.linearize(env, x0, cmd, t0=t0)
    errs = {}
    grad, gn = eng.grad_obj(ex)
    errs["grad_obj"] = relerr(grad[0], g["grad_obj"])
    one = lambda v: torch.tensor([[v]], dtype=torch.float64)
    for reg in (0.0, 0.5):
        tag = f"lq_reg{reg:g}"
        bw = eng.backward(ex, _lib.BWD_LINQUAD, one(reg))
        errs[tag + "_K"] = relerr(bw["K"][0, 0], g[tag + "_K"])
        errs[tag + "_k"] = relerr(bw["k"][0, 0], g[tag + "_k"])
        errs[tag + "_jcst"] = relerr(bw["jcst"][0, 0], g[tag + "_jcst"])
        d, _, _ = eng.rollout(ex, _lib.ROLL_LIN, one(0.7), bw)
        errs[tag + "_rol"] = relerr(d[0, 0], g[tag + "_rol"])
        # reg_ctrl=0 on the cart-pendulum (control cost 1e-6 u^2) gives gains ~1e6 and an explosive exact
        # roll-out in the reference itself: chaotic, not a parity target
        if not (reg == 0.0 and "cart" in name):
            d, _, _ = eng.rollout(ex, _lib.ROLL_EXACT, one(0.7), bw)
            errs[tag + "_roe"] = relerr(d[0, 0], g[tag + "_roe"])
    for mode, mid in (("newton", _lib.BWD_QUAD_NEWTON), ("ddp", _lib.BWD_QUAD_DDP)):
        for reg in (0.5, 20.0):
            tag = f"q_{mode}_reg{reg:g}"
            bw = eng.backward(ex, mid, one(reg))
            errs[tag + "_K"] = relerr(bw["K"][0, 0], g[tag + "_K"])
            errs[tag + "_k"] = relerr(bw["k"][0, 0], g[tag + "_k"])
            errs[tag + "_jcst"] = relerr(bw["jcst"][0, 0], g[tag + "_jcst"])
            assert bool(bw["feasible"][0, 0]) == bool(g[tag + "_feasible"]), (name, tag)
            d, _, _ = eng.rollout(ex, _lib.ROLL_LIN, one(1.0), bw)
            errs[tag + "_rol"] = relerr(d[0, 0], g[tag + "_rol"])
            d, _, _ = eng.rollout(ex, _lib.ROLL_EXACT, one(1.0), bw)
            errs[tag + "_roe"] = relerr(d[0, 0], g[tag + "_roe"])
    bad = {k: v for k, v in errs.items() if not v < TOL}
    assert not bad, (name, bad)
    return errs


# ---- run_min_algo traces ------------------------------------------------------------------------------
def first_stepsize(algo):
    return 100.0 if (algo != "gd" and "reg" in algo.split("_")[-1]) else 1.0


def check_run_teacher_forced(eng, name, max_iters=None):
    """One engine iteration from every iterate (cmd_k, stepsize_k) of the reference trace: new cost and
    new controls within 1e-9 relative, identical number of line-search trips."""
    g = golden(name)
    env, algo = env_of(g), str(g["algo"])
    n = g["cmds"].shape[0] - 1
    if max_iters:
        n = min(n, max_iters)
    x0 = t64(g["x0"]).reshape(1, -1)
    # all iterations as one batch of independent problems
    cmd0 = t64(g["cmds"][:n])
    steps = t64([first_stepsize(algo)] + list(g["stepsize_state"][:n - 1]))
    r = eng.solve(env, algo, 1, x0=x0.expand(n, -1), cmd0=cmd0, stepsize=steps, record_trips=True,
                  handle_bad_dir=bad_dir_of(g))
    out = {}
    for k in range(n):
        e_cost = abs(r.cost[k, 1].item() - g["cost"][k + 1]) / abs(g["cost"][k + 1])
        e_cmd = relerr(r.cmd[k], g["cmds"][k + 1])
        e_c0 = abs(r.cost[k, 0].item() - g["cost"][k]) / abs(g["cost"][k])
        e_step = abs(r.stepsize[k].item() - g["stepsize_state"][k]) / abs(g["stepsize_state"][k])
        out[k] = (e_c0, e_cost, e_cmd, e_step, int(r.ls_trips[k, 0]), int(g["ls_trips"][k]))
        # 'dir' step modes restart every line search from 1 (algos_steps.py:336): the returned step is not
        # optimiser state, and when the search only stops at s ~ 1e-15 its trip count is rounding noise
        if algo.endswith("_dir"):
            e_step = 0.0
        assert e_c0 < TOL and e_cost < TOL and e_cmd < TOL and e_step < TOL, (name, k, out[k])
    return out


def check_run_free(eng, name, tol=TOL):
    """Free-running solve from the reference's initial point: per-iteration cost, final controls."""
    g = golden(name)
    env, algo = env_of(g), str(g["algo"])
    n = g["cmds"].shape[0] - 1
    r = eng.solve(env, algo, n, x0=t64(g["x0"]).reshape(1, -1), record_trips=True, handle_bad_dir=bad_dir_of(g))
    ec = np.abs(r.cost[0].cpu().numpy() - g["cost"]) / np.abs(g["cost"])
    assert np.all(ec < tol), (name, ec)
    assert relerr(r.cmd[0], g["cmds"][-1]) < max(tol, 1e-8), name
    return ec


# ---- generic dense small systems (envs/lin_quad.py, tests/test_auto_lin_quad.py of the reference) ------
def check_lq_synth(eng):
    """lin_quad_backward + roll_out_lin on make_synth_linear_env(H=50, nx=4, nu=4) against the reference's DP
    result AND against its dense autodiff Newton step on the flattened problem (reg_ctrl 0 and 10), for a
    batch of identical problems with the candidates (reg 0, reg 10) evaluated as two line-search slots."""
    g = golden("lq_synth")
    H, nx, nu, B = 50, 4, 4, 3
    rep = lambda a, n: t64(a).unsqueeze(0).unsqueeze(0).expand(B, n, *a.shape).contiguous()
    A, Bm = rep(g["A"], H), rep(g["B"], H)
    Q, q = rep(g["Q"], H), t64(g["grad_ctrl"]).unsqueeze(0).expand(B, H, nu).contiguous()
    P = rep(g["P"], H + 1).clone()
    P[:, 0] = 0.0                                   # costs[0] carries no state cost (forward.py:130-134)
    p = t64(g["grad_state"]).unsqueeze(0).expand(B, H + 1, nx).contiguous()
    reg = t64([[0.0, 10.0]]).expand(B, 2).contiguous()
    K, k, jcst = eng.lq_backward_dense(A, Bm, p, P, q, Q, reg)
    diff = eng.rollout_lin_dense(A, Bm, K, k, torch.ones(B, 2, dtype=torch.float64))
    for c, tag in enumerate(("reg0", "reg10")):
        for b in (0, B - 1):
            assert relerr(K[b, c], g[tag + "_K"]) < TOL and relerr(k[b, c], g[tag + "_k"]) < TOL
            assert abs(float(jcst[b, c]) - float(g[tag + "_jcst"])) < TOL * abs(float(g[tag + "_jcst"]))
            assert relerr(diff[b, c], g[tag + "_diff"]) < TOL
            # dense Newton cross-check of the reference test (it prints 8.5e-10 / 1.3e-11)
            assert relerr(diff[b, c], g[tag + "_newton_diff"]) < 1e-8
            assert abs(float(jcst[b, c]) - float(g[tag + "_newton_opt"])) < 1e-8 * abs(float(g[tag + "_newton_opt"]))


# ---- randomised configurations against the oracle (beyond the golden vectors) ---------------------------
def random_env_cfg(seed):
    """A random but valid environment configuration: every constructor argument the models take is drawn
    (integrator, cost, time penalty, track, weights, dt, horizon offsets, barrier weights, physical times)."""
    r = np.random.default_rng(seed)
    kind = ["pendulum", "cart_pendulum", "simple_car", "real_car"][seed % 4]
    u = lambda lo, hi: float(r.uniform(lo, hi))
    if kind == "pendulum":
        return dict(env="pendulum", horizon=4, dt=u(0.01, 0.06), reg_speed=u(0.05, 0.5), reg_ctrl=u(1e-6, 1e-2),
                    stay_put_time=[None, 0.05][int(r.integers(2))])
    if kind == "cart_pendulum":
        return dict(env="cart_pendulum", horizon=4, dt=u(0.02, 0.06), reg_speed=u(0.05, 0.5), reg_ctrl=u(1e-6, 1e-2),
                    reg_barrier=u(0.5, 5.0), stay_put_time=u(0.03, 0.12), x_limits=[None, (-0.4, 0.4)][int(r.integers(2))],
                    discretization=["euler", "rk4cst", "rk4"][int(r.integers(3))])
    cfg = dict(env=kind, horizon=3, dt=u(0.01, 0.05), track=["simple", "complex", "circle", "line", "bend"][int(r.integers(5))],
               discretization=["euler", "rk4_cst_ctrl", "rk4"][int(r.integers(3))], cost=["contouring", "exact"][int(r.integers(4) == 0)],
               time_penalty=["vref_squared", "tref", "dref", "dref_squared"][int(r.integers(4))],
               reg_cont=u(0.05, 5.0), reg_lag=u(0.5, 20.0), reg_speed=u(0.05, 2.0), reg_bar=u(0.0, 200.0), vref=u(1.0, 3.5),
               vinit=u(0.6, 2.0), reg_ctrl=[u(1e-6, 1e-3), (u(1e-6, 1e-3), u(1e-6, 1e-3))][int(r.integers(2))])
    if int(r.integers(3)) == 0:
        cfg["reg_obs"] = u(0.5, 50.0)
    return cfg


def check_random_config_vs_oracle(eng, seed, tol=1e-9):
    """Expansion (first and second order), backward passes (linquad / Newton / DDP) and both roll-outs of a random
    configuration at a random state and controls: engine vs the oracle's autograd (oracle/ilqc_oracle.py)."""
    from oracle import ilqc_oracle as orc
    cfg = random_env_cfg(seed)
    env, oenv = make_env(cfg), orc.make_env(cfg)
    r = np.random.default_rng(1000 + seed)
    nx, nu, H = env.dim_state, env.dim_ctrl, env.horizon
    x0 = oenv.init_state.detach().clone()
    assert relerr(env.init_state, x0.numpy()) < 1e-14
    x0 = x0 + torch.tensor(0.05 * r.normal(size=nx))
    if cfg["env"] in ("simple_car", "real_car"):
        x0[-1] = abs(float(x0[-1])) + 0.3          # vtref > 0 (log barrier)
        x0[-2] = abs(float(x0[-2])) + float(r.uniform(0.0, 3.0))   # somewhere along the track
    cmd = torch.tensor(0.3 * r.normal(size=(H, nu)))
    oenv.init_state = x0.clone()
    traj, costs, steps = orc.forward(oenv, cmd.clone().requires_grad_(True), approx="quad")
    lam = torch.tensor(r.normal(size=(H, nx)))
    ex = eng.expand_dense(env, x0.reshape(1, -1), cmd.unsqueeze(0), lam=lam.unsqueeze(0))
    tr, ct, tot = eng.rollout_cost(env, x0.reshape(1, -1), cmd.unsqueeze(0))
    errs = dict(traj=relerr(tr[0], torch.stack(traj).detach().numpy()),
                total=abs(float(tot[0]) - float(sum(costs))) / (abs(float(sum(costs))) + 1e-300))
    for t, st in enumerate(steps):
        hxx, huu, hxu = orc.hess_dyn(st, lam[t], nx, nu)
        for key, a, b in (("A", ex["A"][0, t], st.A), ("B", ex["B"][0, t], st.B), ("p", ex["p"][0, t + 1], st.p),
                          ("P", ex["P"][0, t + 1], st.P), ("q", ex["q"][0, t], st.q.reshape(nu)),
                          ("Q", ex["Q"][0, t], st.Q.reshape(nu, nu)), ("Huu", ex["Huu"][0, t], huu), ("Hxu", ex["Hxu"][0, t], hxu)):
            errs[f"{key}{t}"] = float((a.cpu() - b.detach()).abs().max() / (b.detach().abs().max() + 1e-12))
        if t > 0:
            errs[f"Hxx{t}"] = float((ex["Hxx"][0, t].cpu() - hxx.detach()).abs().max() / (hxx.detach().abs().max() + 1e-12))
    lin = eng.linearize(env, x0.reshape(1, -1), cmd.unsqueeze(0))
    one = lambda v: torch.tensor([[v]], dtype=torch.float64)
    for mode, mid in (("linquad", _lib.BWD_LINQUAD), ("newton", _lib.BWD_QUAD_NEWTON), ("ddp", _lib.BWD_QUAD_DDP)):
        reg = 0.7
        gains, jcst, feas = orc.backward(steps, reg, mode=mode)
        bw = eng.backward(lin, mid, one(reg))
        errs[mode + "_K"] = relerr(bw["K"][0, 0], torch.stack([K for K, _ in gains]).detach().numpy())
        errs[mode + "_k"] = relerr(bw["k"][0, 0], torch.stack([k for _, k in gains]).detach().numpy())
        errs[mode + "_jcst"] = abs(float(bw["jcst"][0, 0]) - float(jcst)) / (abs(float(jcst)) + 1e-300)
        d, _, _ = eng.rollout(lin, _lib.ROLL_LIN, one(0.6), bw)
        errs[mode + "_rol"] = relerr(d[0, 0], orc.roll_out_lin(steps, gains, 0.6).detach().numpy())
        d, _, _ = eng.rollout(lin, _lib.ROLL_EXACT, one(0.6), bw)
        errs[mode + "_roe"] = relerr(d[0, 0], orc.roll_out_exact(oenv, traj, gains, cmd, 0.6).detach().numpy())
    # two solver iterations from these controls with a second- and a first-order algorithm (the solver contracts the
    # stored dynamics-Hessian tensor, eng.backward above recomputes it: both paths are compared with the oracle)
    for algo in (["ddp_quad_reg", "classic_quad_reg"][seed % 2], ["ddp_linquad_reg", "classic_linquad_dir"][(seed // 2) % 2]):
        oenv.init_state = x0.clone()
        ocmd, _, m = orc.run_min_algo(oenv, max_iter=2, algo=algo, prev_cmd=cmd.clone())
        res = eng.solve(env, algo, 2, x0=x0.reshape(1, -1), cmd0=cmd.unsqueeze(0))
        n = min(len(m["cost"]), int(res.n_done[0]) + 1)
        errs[algo + "_cost"] = relerr(res.cost[0, :n], np.array(m["cost"][:n]))
        if ocmd is not None and n == 3:
            errs[algo + "_cmd"] = relerr(res.cmd[0], ocmd.detach().numpy())
    bad = {k: v for k, v in errs.items() if not v < tol}
    assert not bad, (seed, cfg, bad)
    return errs


# ---- dense cross-check at batch scale (the idea of the reference's tests/test_auto_lin_quad.py, tests/utils.py) -------
def check_riccati_vs_dense_newton(eng, B=48, H=10, reg=0.3):
    """lin_quad_backward + roll_out_lin is the Newton step of the linear-quadratic model.  For a batch of random
    bicycle problems the step from the Riccati kernels is compared with a dense solve of the same model assembled
    in torch from the dense expansion: dx = G v (block lower-triangular from A_t, B_t), Hessian (Q + reg) + G^T P G,
    gradient q + G^T p, v* = -Hess^{-1} grad."""
    env = make_env(dict(env="real_car", track="simple", cost="contouring", discretization="rk4_cst_ctrl", horizon=H, dt=0.02))
    g = torch.Generator().manual_seed(7)
    nx, nu = env.dim_state, env.dim_ctrl
    x0 = env.init_state.reshape(1, -1).repeat(B, 1)
    x0[:, 3] = 1.0 + 0.05 * torch.randn(B, generator=g, dtype=torch.float64)
    x0[:, 7] = x0[:, 3]
    cmd = 0.2 * torch.randn(B, H, nu, generator=g, dtype=torch.float64)
    ex = eng.expand_dense(env, x0, cmd)
    A, Bm, p, P, q, Q = (ex[k].cpu() for k in ("A", "B", "p", "P", "q", "Q"))
    n = H * nu
    G = torch.zeros(B, H + 1, nx, n, dtype=torch.float64)      # dx_t = G[t] v
    for t in range(H):
        G[:, t + 1] = A[:, t] @ G[:, t]
        G[:, t + 1, :, t * nu:(t + 1) * nu] += Bm[:, t]
    Hess = torch.zeros(B, n, n, dtype=torch.float64)
    grad = q.reshape(B, n).clone()
    for t in range(H):
        Hess[:, t * nu:(t + 1) * nu, t * nu:(t + 1) * nu] += Q[:, t] + reg * torch.eye(nu, dtype=torch.float64)
    for t in range(1, H + 1):
        Hess += G[:, t].transpose(1, 2) @ P[:, t] @ G[:, t]
        grad += (G[:, t].transpose(1, 2) @ p[:, t].unsqueeze(-1)).squeeze(-1)
    v_dense = -torch.linalg.solve(Hess, grad.unsqueeze(-1)).squeeze(-1).reshape(B, H, nu)
    lin = eng.linearize(env, x0, cmd)
    bw = eng.backward(lin, _lib.BWD_LINQUAD, torch.full((B, 1), reg, dtype=torch.float64))
    d, _, _ = eng.rollout(lin, _lib.ROLL_LIN, torch.ones(B, 1, dtype=torch.float64), bw)
    err = relerr(d[:, 0], v_dense.numpy())
    # model decrease: jcst = 1/2 grad . v*
    dec = 0.5 * (grad * v_dense.reshape(B, n)).sum(1)
    err_j = relerr(bw["jcst"][:, 0], dec.numpy())
    assert err < 1e-8 and err_j < 1e-8, (err, err_j)
    return err, err_j


def check_time_parallel_expansion(eng):
    """ilqc_algo_desc.expansion = 1 (one thread per (problem, time step)) against the default fused kernel: same
    mathematics, rounding-level differences only; and against the reference trace of the bicycle H=25 run."""
    g = golden("run_bcar_h25_ddp_linquad_reg")
    env = env_of(g)
    x0 = t64(g["x0"]).reshape(1, -1).repeat(5, 1)
    for algo in ("ddp_linquad_reg", "ddp_quad_reg", "classic_linquad_dir"):
        a = eng.solve(env, algo, 3, x0=x0, record_trips=True)
        b = eng.solve(env, algo, 3, x0=x0, record_trips=True, expansion=1)
        assert relerr(b.cost, a.cost.cpu().numpy()) < 1e-10 and relerr(b.cmd, a.cmd.cpu().numpy()) < 1e-8, algo
        assert torch.equal(a.ls_trips, b.ls_trips) and torch.equal(a.status, b.status), algo
    b = eng.solve(env, "ddp_linquad_reg", 3, x0=x0, expansion=1)
    assert relerr(b.cost[0], g["cost"][:4]) < TOL and relerr(b.cmd[0], g["cmds"][3]) < TOL
