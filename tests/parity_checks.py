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
    # metrics rows of collect_info (run_min_algo.py:156-161): ||grad_u f|| (relative to the first norm: converged runs end
    # at the rounding floor) and the reported step (rescaled by the previous gradient norm in 'reg' mode)
    gn, gref = r.norm_grad_obj[0].cpu().numpy(), g["norm_grad_obj"]
    assert np.all(np.abs(gn - gref) <= 1e-7 * np.abs(gref) + 1e-11 * np.abs(gref[0])), (name, gn, gref)
    sr, sref = r.stepsize_reported[0].cpu().numpy(), g["stepsize_reported"]
    if not algo.endswith("_dir"):     # ('dir' searches that only stop at s ~ 1e-15 report rounding noise)
        # (the reported step of the 'reg' mode is the step state times the previous ||grad_u f||, and the state itself is
        #  rescaled by gradient norms every iteration: it inherits the absolute floor 1e-11 ||grad_u f(u_0)|| of the norms
        #  above as a relative error, accumulated along the chain - factor 4)
        floor = np.concatenate(([0.0], 4e-11 * np.abs(gref[0]) / np.abs(gref[:-1])))
        assert np.all(np.abs(sr - sref) <= (1e-7 + floor) * np.abs(sref)), (name, sr, sref)
    assert np.array_equal(r.ls_trips[0].cpu().numpy()[:n], g["ls_trips"][:n]) or algo.endswith("_dir"), (name, r.ls_trips[0], g["ls_trips"])
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
    # state regularisation (reg_state of lin_quad_backward / quad_backward, backward.py:35, :105): added to EVERY diagonal
    # entry of the state-cost Hessian, also outside the model's stored sparsity pattern
    for mode, mid in (("linquad", _lib.BWD_LINQUAD), ("ddp", _lib.BWD_QUAD_DDP)):
        gains, jcst, feas = orc.backward(steps, 0.7, reg_state=0.3, mode=mode)
        bw = eng.backward(lin, mid, one(0.7), reg_state=0.3)
        errs[mode + "_regstate_K"] = relerr(bw["K"][0, 0], torch.stack([K for K, _ in gains]).detach().numpy())
        errs[mode + "_regstate_jcst"] = abs(float(bw["jcst"][0, 0]) - float(jcst)) / (abs(float(jcst)) + 1e-300)
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


# ---- batched reference traces at SURVEY 8(d)'s sampling plan (tests/golden/make_golden_batch.py) -----------------
def batch_inputs(g):
    """x0 [P,nx], weights [P,3] or None, env of a batch_<group>.npz fixture."""
    env = env_of(g)
    x0 = t64(g["x0"])
    w = t64(g["weights"]) if "weights" in g.files else None
    return env, x0, w


SENS_METRICS = ("cost1", "cmd", "gn0", "gn1", "step", "srep")


def sensitivity_rows(name):
    """{(p, k): ({metric: measured movement of the reference's own output under 1e-15 input noise}, trips_unstable)} for
    the rows of fixture `name` listed in tests/golden/sensitivity.npz (tests/golden/make_sensitivity.py)."""
    path = os.path.join(GOLDEN, "sensitivity.npz")
    if not os.path.exists(path):
        return {}
    z = np.load(path)
    out = {}
    for i in range(len(z["p"])):
        if str(z["fixture"][i]) == name:
            out[(int(z["p"][i]), int(z["k"][i]))] = ({m: float(z["sens_" + m][i]) for m in SENS_METRICS}, int(z["trips_unstable"][i]))
    return out


def batch_teacher_forced_errors(eng, name, **solve_kw):
    """EVERY stored iterate of EVERY problem of the fixture as one batch of independent one-iteration solves (teacher
    forcing).  Returns ({metric: relative error per row}, trips equal per row, rows = [(problem, iteration)])."""
    g = golden(name)
    env, x0, w = batch_inputs(g)
    algo = str(g["algo"])
    P = x0.shape[0]
    rows = [(p, k) for p in range(P) for k in range(int(g["n_done"][p]))]
    pi = np.array([r[0] for r in rows]); ki = np.array([r[1] for r in rows])
    cmd0 = t64(g["cmds"][pi, ki])
    s_prev = np.where(ki > 0, g["stepsize_state"][pi, np.maximum(ki - 1, 0)], first_stepsize(algo))
    r = eng.solve(env, algo, 1, x0=x0[pi], cmd0=cmd0, stepsize=t64(s_prev), weights=None if w is None else w[pi],
                  record_trips=True, **solve_kw)
    rel = lambda a, b: np.abs(a - b) / np.abs(b)
    cost = r.cost.cpu().numpy(); gn = r.norm_grad_obj.cpu().numpy(); srep = r.stepsize_reported.cpu().numpy()
    e = dict(cost0=rel(cost[:, 0], g["cost"][pi, ki]), cost1=rel(cost[:, 1], g["cost"][pi, ki + 1]),
             gn0=rel(gn[:, 0], g["norm_grad_obj"][pi, ki]), gn1=rel(gn[:, 1], g["norm_grad_obj"][pi, ki + 1]),
             step=rel(r.stepsize.cpu().numpy(), g["stepsize_state"][pi, ki]),
             srep=rel(srep[:, 1], g["stepsize_reported"][pi, ki + 1]))
    cmd = r.cmd.cpu().numpy()
    ref = g["cmds"][pi, ki + 1]
    e["cmd"] = np.abs(cmd - ref).reshape(len(rows), -1).max(1) / np.abs(ref).reshape(len(rows), -1).max(1)
    trips_ok = r.ls_trips[:, 0].cpu().numpy() == g["ls_trips"][pi, ki]
    return e, trips_ok, rows


def check_batch_teacher_forced(eng, name, tol=TOL, **solve_kw):
    """Teacher-forced parity of every stored iterate: new cost, new controls, step-size state, gradient norms and reported
    step within `tol` relative, identical line-search trip counts - except on the rows where the REFERENCE's own result
    moves by more than that when its input is perturbed in the last bit (tests/golden/make_sensitivity.py measures it with
    the unmodified reference): there the tolerance of a metric is 4 x the reference's measured movement of that metric, and
    a row on which the reference's own trip count flips under the perturbation is not compared at all.  At most 2 % of
    the rows may have a relaxed COST tolerance or be skipped, not counting iterations at numerical convergence (cost
    change below the parity tolerance: the cart-pendulum fixture spends its iterations 13-19 there and the reference's
    own trip count flips on 230 of those 450 rows).  Returns the worst errors."""
    e, trips_ok, rows = batch_teacher_forced_errors(eng, name, **solve_kw)
    n = len(rows)
    dump = os.environ.get("ILQC_DUMP_ROWS")
    if dump:   # rows for make_sensitivity.py (the CUDA engine's rounding differs from the host simulator's)
        import json
        os.makedirs(dump, exist_ok=True)
        flagged = [list(map(int, rows[i])) for i in range(n) if not trips_ok[i] or any(e[k][i] > 1e-10 for k in SENS_METRICS)]
        with open(os.path.join(dump, name + ".json"), "w") as f:
            json.dump(dict(rows=flagged, device=str(getattr(eng, "device", "?"))), f)
    tols = {k: np.full(n, tol) for k in e}
    skip = np.zeros(n, bool)
    sens = sensitivity_rows(name)
    gc = golden(name)["cost"]
    n_relaxed = n_conv_skipped = 0
    for i, row in enumerate(rows):
        if tuple(row) in sens:
            sm, unstable = sens[tuple(row)]
            # an iteration that changes the cost by less than the parity tolerance itself: the iterate has converged
            # numerically and the reference's line search there is decided by rounding (measured: make_sensitivity.py)
            converged = (gc[row[0], row[1]] - gc[row[0], row[1] + 1]) < tol * abs(gc[row[0], row[1]])
            if unstable:
                skip[i] = True
                n_relaxed += int(not converged)
                n_conv_skipped += int(converged)
                continue
            for k in SENS_METRICS:
                tols[k][i] = max(tol, 4 * sm[k])
            n_relaxed += int(tols["cost1"][i] > tol and not converged)
    # before numerical convergence at most 2 % of the rows may be skipped or have a relaxed COST tolerance
    assert n_relaxed <= max(1, n // 50), (name, n_relaxed)
    pi = np.array([r[0] for r in rows]); ki = np.array([r[1] for r in rows])
    conv_row = (gc[pi, ki] - gc[pi, ki + 1]) < tol * np.abs(gc[pi, ki])
    # gradient norms have an absolute floor (cancellation in the adjoint sum): 1e-11 of the problem's largest norm; the
    # reported 'reg' step is the step state times the previous norm and inherits it as a relative error
    gng = golden(name)["norm_grad_obj"]
    gmax = np.nanmax(gng, axis=1)[pi]
    ok = {k: e[k] < tols[k] for k in e}
    ok["gn0"] |= e["gn0"] * np.abs(gng[pi, ki]) <= 1e-11 * gmax
    ok["gn1"] |= e["gn1"] * np.abs(gng[pi, ki + 1]) <= 1e-11 * gmax
    ok["srep"] |= e["srep"] <= 4e-11 * gmax / np.abs(gng[pi, ki])
    # on a numerically converged row the search is decided by the last bits of two equal costs: when the trip count
    # differs there, only the costs (and the gradient norm of the given iterate) are compared
    flipped = conv_row & ~trips_ok
    for k in ("cmd", "gn1", "step", "srep"):
        ok[k] |= flipped
    worst = {k: float(np.max(v[~skip])) for k, v in e.items()}
    bad = {k: [(float(e[k][i]), rows[i]) for i in np.nonzero(~ok[k] & ~skip)[0][:4]] for k in e}
    bad = {k: v for k, v in bad.items() if v}
    assert not bad, (name, bad)
    mism = np.nonzero(~trips_ok & ~skip & ~conv_row)[0]
    assert len(mism) == 0, (name, [rows[i] for i in mism[:8]])
    worst["n_rows"] = n
    worst["n_relaxed_cost_or_skipped"] = n_relaxed
    worst["n_converged_rows_skipped"] = n_conv_skipped
    worst["n_converged_rows_with_other_trip_count"] = int(flipped.sum())
    worst["n_rows_with_measured_sensitivity"] = len(sens)
    return worst


def check_batch_free(eng, name, tol=TOL, min_match=3, **solve_kw):
    """Free-running solve of every problem of the fixture from the reference's starting point.  Statistic of SURVEY
    section 4: the first iteration at which a problem's cost leaves the reference's by more than `tol` (free chains
    amplify rounding once the iterates are converged; a flipped line-search trip changes the rest).  Asserts that every
    problem matches the reference on its first `min_match` iterations and returns the per-problem first diverging
    iteration (n+1 = never)."""
    g = golden(name)
    env, x0, w = batch_inputs(g)
    algo = str(g["algo"])
    n = g["cmds"].shape[1] - 1
    r = eng.solve(env, algo, n, x0=x0, weights=w, record_trips=True, **solve_kw)
    cost = r.cost.cpu().numpy()
    ref = g["cost"]
    err = np.abs(cost - ref) / np.abs(ref)
    err = np.where(np.isnan(ref) & np.isnan(cost), 0.0, err)       # both stopped
    err = np.where(np.isnan(err), np.inf, err)
    first = np.array([next((k for k in range(n + 1) if not err[p, k] < tol), n + 1) for p in range(err.shape[0])])
    min_match = min(min_match, n - 1)     # (the ddp_quad_reg fixtures hold 2 iterations: iteration 1 has to match; the second
    # one starts from an iterate that differs by rounding and amplifies it - the teacher-forced test covers it at 1e-9)
    assert (first > min_match).all(), (name, first, err[:, :min_match + 1].max())
    return first


# ---- MPC windows of the reference (chained mpc_step, perturbed episodes and the variants) ------------------------
def mpc_optim(g):
    o = ast.literal_eval(str(g["optim"]))
    kw = {k: o[k] for k in ("optim_on_full_window", "keep_applying_last_ctrl") if k in o}
    return o, kw


def check_mpc_teacher_forced(name, n_strict=6):
    """Drop-in mpc_step (default engine) on every window of a reference episode, warm-started from the REFERENCE's
    previous window: re-anchored state / time index, per-iteration costs (strict on the first iterations, see
    tests/test_host_logic.py::_check_window for the tail), spliced controls."""
    from ilqc_b200.model_predictive_control import mpc_step
    g = golden(name)
    o, kw = mpc_optim(g)
    cfg = cfg_of(g)
    out = []
    for w in range(o["windows"]):
        env = make_env(cfg)
        env.init_state = t64(g["x0"])
        prev = None if w == 0 else t64(g[f"cmd_opt_{w - 1}"])
        cmd_opt, m = mpc_step(env, full_horizon=int(g[f"full_horizon_{w}"]), overlap=o["overlap"], max_iter=o["max_iter"],
                              algo="ddp_linquad_reg", prev_cmd=prev, **kw)
        assert relerr(env.init_state, g[f"x0_{w}"]) < TOL and env.init_time_iter == int(g[f"t0_{w}"]), (name, w)
        c, cref = np.array(m["cost"]), g[f"cost_{w}"]
        k = min(len(c), len(cref), n_strict)
        assert k >= min(n_strict, len(cref)) and np.max(np.abs(c[:k] - cref[:k]) / np.abs(cref[:k])) < TOL, (name, w, c, cref)
        n = min(len(c), len(cref))
        assert np.max(np.abs(c[:n] - cref[:n]) / np.abs(cref[:n])) < 1e-3, (name, w)
        assert cmd_opt.shape == g[f"cmd_opt_{w}"].shape, (name, w)
        # the first iterations are also compared through their reported gradient norms and steps
        gnr, sr = g[f"norm_grad_obj_{w}"], g[f"stepsize_{w}"]
        # (warm-started windows start near a stationary point: their gradient norms sit at the cancellation floor of the
        #  adjoint sum, 1e-10 of the cold-start norm of the episode's first window)
        gfloor = 1e-10 * float(g["norm_grad_obj_0"][0])
        assert np.all(np.abs(np.array(m["norm_grad_obj"][:k]) - gnr[:k]) <= 1e-7 * np.abs(gnr[:k]) + gfloor), (name, w)
        # (a line search that only ends after 25+ halvings sits at a converged point: whether s or s/2 passes the test
        #  `new < curr + dec` is decided by the last bit of the cost, so those steps are not compared)
        trips = g[f"trips_{w}"]
        ok = np.array([j == 0 or trips[j - 1] < 25 for j in range(k)])
        sfloor = np.concatenate(([0.0], 4 * gfloor / np.abs(gnr[:k - 1]))) if k > 1 else np.zeros(k)
        assert np.all((np.abs(np.array(m["stepsize"][:k]) - sr[:k]) <= (1e-7 + sfloor) * np.abs(sr[:k]))[ok]), (name, w)
        out.append((len(c), len(cref)))
    return out


def check_mpc_batched_vs_reference(eng, names, n_strict=6):
    """run_mpc_batched (ilqc_mpc, the window loop on the device side) on the reference's episodes as ONE batch:
    free-running, so the first window is compared strictly (1e-9 on its first iterations); later windows start from the
    previous window's own result, whose near-converged tail has drifted from the reference's (test_host_logic.py
    ::check_window: 1e-3 on costs, 5e-2 on controls), so only a sanity bound (20 %) is put on their starting cost: the chain of near-converged windows is chaotic (measured: up to 10 % by window 3).  The per-window
    parity proper is the teacher-forced check above."""
    from ilqc_b200.model_predictive_control import run_mpc_batched
    gs = [golden(n) for n in names]
    o, kw = mpc_optim(gs[0])
    env = make_env(dict(cfg_of(gs[0]), reg_cont=1.0, reg_lag=1.0, reg_speed=1.0))
    x0 = t64(np.stack([g["x0"] for g in gs]))
    w = t64(np.stack([g["weights"] for g in gs]))
    fh = int(gs[0][f"full_horizon_{o['windows'] - 1}"])
    r = run_mpc_batched(env, full_horizon=fh, window_size=o["window_size"], overlap=o["overlap"], max_iter=o["max_iter"],
                        x0=x0, weights=w, engine=eng, **kw)
    assert r.horizons == [int(gs[0][f"full_horizon_{k}"]) for k in range(o["windows"])]
    for i, g in enumerate(gs):
        c0 = g["cost_0"]
        k = min(n_strict, len(c0))
        assert relerr(r.cost[i, 0, :k], c0[:k]) < TOL, (names[i], r.cost[i, 0], c0)
        for wdw in range(1, o["windows"]):
            cw = g[f"cost_{wdw}"]
            assert abs(float(r.cost[i, wdw, 0]) - cw[0]) / abs(cw[0]) < 0.2, (names[i], wdw, float(r.cost[i, wdw, 0]), cw[0])
        assert r.cmd.shape[1] == g[f"cmd_opt_{o['windows'] - 1}"].shape[0]
    return r


# ---- outcomes that end the loop by status (run_min_algo.py:195-264) -----------------------------------------------
def check_status_cases(eng):
    from ilqc_b200.algorithms.run_min_algo import run_min_algo, check_cvg_status
    g = golden("status_cases")
    res = {}
    for tag in ("nan0", "stuck0", "div", "nanstep", "conv"):
        cfg = ast.literal_eval(str(g[tag + "_cfg"]))
        env = make_env(cfg)
        env.init_state = t64(g[tag + "_x0"])
        kw = dict(max_iter=int(g[tag + "_max_iter"]), algo=str(g[tag + "_algo"]))
        if tag + "_stepsize" in g.files and tag in ("stuck0", "div", "nanstep"):
            kw["stepsize"] = float(np.atleast_1d(g[tag + "_stepsize"])[0])
        if tag + "_line_search" in g.files:
            kw["line_search"] = bool(g[tag + "_line_search"])
        if tag + "_cmd0" in g.files:
            kw["prev_cmd"] = t64(g[tag + "_cmd0"])
        cmd, aux, m = run_min_algo(env, **kw)
        cref = np.atleast_1d(g[tag + "_cost"])
        c = np.array(m["cost"])
        if tag == "conv":
            # 'converged' = two consecutive costs equal to 1e-24 relative, i.e. bit-identical (run_min_algo.py:218-220):
            # WHICH iteration first repeats the previous cost bit for bit is decided by the last rounding (the FMA-contracted
            # device arithmetic may get there one iteration before / after the reference) - the rows both have must agree
            assert abs(len(c) - len(cref)) <= 1 and len(c) >= 3, (tag, c, cref)
            n = min(len(c), len(cref))
            c, cref = c[:n], cref[:n]
            assert check_cvg_status(m, verbose=False) == str(g[tag + "_status"]) == "converged"
            m = {k: list(v)[:n] for k, v in m.items()}
        assert len(c) == len(cref), (tag, c, cref)
        both_nan = np.isnan(c) & np.isnan(cref)
        assert np.all(both_nan | (np.abs(c - cref) <= TOL * np.abs(cref))), (tag, c, cref)
        assert tag == "conv" or check_cvg_status(m, verbose=False) == str(g[tag + "_status"]), (tag, check_cvg_status(m, verbose=False))
        gr = np.atleast_1d(g[tag + "_norm_grad_obj"])[:len(c)]
        # (relative to the first gradient norm: converged runs end with norms at the rounding floor)
        ok = np.isnan(gr) | (np.abs(np.array(m["norm_grad_obj"]) - gr) <= 1e-7 * np.abs(gr) + 1e-12 * np.abs(gr[0]))
        assert ok.all(), (tag, m["norm_grad_obj"], gr)
        res[tag] = (len(c), str(g[tag + "_status"]))
    # the same problems through the batched engine: status words
    return res


# ---- lower-level reference API: algos_steps ---------------------------------------------------------------------------
def check_min_step_chain(name, iters):
    """min_step (algos_steps.py:277-381) chained by hand like run_min_algo's loop, against the reference trace."""
    from ilqc_b200.algorithms import min_step
    from ilqc_b200.algorithms.run_min_algo import check_nomenclature
    g = golden(name)
    env, algo = env_of(g), str(g["algo"])
    env.init_state = t64(g["x0"])
    _, approx, _ = check_nomenclature(algo)
    cmd = t64(g["cmds"][0])
    stepsize = first_stepsize(algo)
    traj, costs = env.forward(cmd, approx=approx)
    for k in range(iters):
        cmd, traj, costs, stepsize = min_step(env, traj, costs, cmd, stepsize, True, algo, "check_total_value")
        assert relerr(cmd, g["cmds"][k + 1]) < TOL, (name, k)
        assert abs(float(sum(costs)) - g["cost"][k + 1]) < TOL * abs(g["cost"][k + 1]), (name, k)
        if not algo.endswith("_dir"):
            assert abs(float(stepsize) - g["stepsize_state"][k]) < TOL * abs(g["stepsize_state"][k]), (name, k)


def check_newton_oracle():
    """tests/test_newton.py of the reference with only its import lines changed: the structured Newton direction
    (classic_oracle, Riccati kernels) against the dense Newton step on the flattened problem (newton_oracle,
    ilqc_newton_dense); both against the reference's own vectors (newton_pendulum.npz)."""
    from ilqc_b200.algorithms.algos_steps import newton_oracle, classic_oracle
    from ilqc_b200.envs.pendulum import Pendulum
    g = golden("newton_pendulum")
    env = Pendulum(horizon=100, stay_put_time=1.)
    cmd = t64(g["cmd"])
    traj, costs = env.forward(cmd, approx='quad')
    stepsize = 1.
    ilqc_oracle = classic_oracle(traj, costs, approx='quad', step_mode='dir', handle_bad_dir='check_total_value')
    auto_diff_oracle = newton_oracle(env, cmd)
    cmd_ilqc, dec_ilqc = ilqc_oracle(stepsize)
    cmd_newton, dec_newton = auto_diff_oracle(stepsize)
    assert float(torch.linalg.norm(cmd_ilqc - cmd_newton)) < 1e-8 * float(torch.linalg.norm(cmd_newton))
    assert relerr(cmd_ilqc, g["ilqc_dir"]) < TOL and relerr(cmd_newton, g["autodiff_dir"]) < 1e-8
    assert abs(float(dec_newton) - float(g["autodiff_dec"])) < 1e-8 * abs(float(g["autodiff_dec"]))
    assert abs(float(dec_ilqc) - float(g["ilqc_dec"])) < TOL * abs(float(g["ilqc_dec"]))


def check_incremental_and_restart():
    """solve_ctrl_pb_incrementally (fhc_pipeline.py:46-53): 10 iterations per leg through the restart arguments ==
    one straight run; the restarted legs test divergence against the ORIGINAL first cost (ilqc_algo_desc.pp_cost0)."""
    from ilqc_b200.finite_horizon_control import solve_ctrl_pb_incrementally, solve_ctrl_pb
    from ilqc_b200.algorithms.run_min_algo import run_min_algo, check_cvg_status
    g = golden("run_pend_h100_ddp_linquad_reg")
    cfg = cfg_of(g)
    out = solve_ctrl_pb_incrementally(cfg, dict(max_iter=20, algo="ddp_linquad_reg"))
    m = out["metrics"]
    assert m["iteration"] == list(range(21)) and relerr(torch.tensor(m["cost"]), g["cost"]) < TOL
    assert relerr(out["cmd_opt"], g["cmds"][20]) < TOL
    assert relerr(torch.tensor(m["norm_grad_obj"]), g["norm_grad_obj"]) < 1e-6
    straight = solve_ctrl_pb(cfg, dict(max_iter=20, algo="ddp_linquad_reg"))
    assert relerr(torch.tensor(m["cost"]), np.array(straight["metrics"]["cost"])) < 1e-12
    # restart + divergence: the fixed-step gradient run of status_cases 'div' stops after one iteration because the cost
    # exceeds cost[0] + 1e3; restarted from an (artificial) history whose first cost is huge it must keep going
    gs = golden("status_cases")
    env = make_env(ast.literal_eval(str(gs["div_cfg"])))
    kw = dict(algo="gd", stepsize=3e3, line_search=False)
    cmd0 = t64(gs["div_cmd0"])
    _, aux, m0 = run_min_algo(env, max_iter=0, prev_cmd=cmd0.clone(), **kw)
    _, _, m1 = run_min_algo(env, max_iter=3, prev_cmd=cmd0.clone(), optim_aux_vars=aux, past_metrics={k: list(v) for k, v in m0.items()}, **kw)
    assert len(m1["cost"]) == 2 and check_cvg_status(m1, verbose=False) == "diverged"
    fake = {k: list(v) for k, v in m0.items()}
    fake["cost"][0] = 1e9
    _, _, m2 = run_min_algo(env, max_iter=3, prev_cmd=cmd0.clone(), optim_aux_vars=aux, past_metrics=fake, **kw)
    assert len(m2["cost"]) > 2 and abs(m2["cost"][1] - m1["cost"][1]) < 1e-12 * abs(m1["cost"][1])


def check_dense_diagnostics():
    """compute_min_sing_eigval_jac's Jacobian (run_min_algo.py:103-124) and a newton_dir run (algos_steps.py:46-102
    through run_min_algo) against the reference (tests/golden/dense_cases.npz)."""
    from ilqc_b200.algorithms import run_min_algo, compute_min_sing_eigval_jac
    g = golden("dense_cases")
    for tag in ("pend", "bcar"):
        env = make_env(ast.literal_eval(str(g[tag + "_cfg"])))
        cmd = t64(g[tag + "_cmd"])
        jac = env._engine().traj_jacobian(env, env.init_state.reshape(1, -1), cmd.unsqueeze(0), t0=env.init_time_iter)[0].cpu()
        ref = t64(g[tag + "_jac_t"]).t()
        assert jac.shape == ref.shape and float((jac - ref).abs().max()) < TOL * float(ref.abs().max()), tag
        # the reference's number is the smallest singular value of a rank-deficient Gram matrix: rounding noise
        sev, sref = compute_min_sing_eigval_jac(cmd, env), float(g[tag + "_min_sev"])
        smax = float(torch.linalg.svdvals(ref @ ref.t()).max())
        assert abs(sev - sref) < 1e-12 * smax, (tag, sev, sref)
    env = make_env(ast.literal_eval(str(g["newton_cfg"])))
    cmd, aux, m = run_min_algo(env, max_iter=5, algo="newton_dir", prev_cmd=t64(g["newton_cmd0"]))
    assert relerr(torch.tensor(m["cost"]), g["newton_cost"]) < 1e-8, (m["cost"], g["newton_cost"])
    assert relerr(cmd, g["newton_cmd"]) < 1e-7
    assert relerr(torch.tensor(m["norm_grad_obj"]), g["newton_norm_grad_obj"]) < 1e-6
    assert relerr(torch.tensor([float(s) for s in m["stepsize"]]), g["newton_stepsize"]) < TOL
