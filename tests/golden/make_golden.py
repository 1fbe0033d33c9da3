"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py [group ...]      # groups: tracks lin lq steps runs mpc cached

The reference is imported from /root/reference with stub ``matplotlib``/``seaborn`` modules
(envs/car.py:7-9 imports matplotlib at module import; it is the only missing import on the path).
Every array is float64.  Nothing from the reference is copied: only its *outputs* are stored.

Fixture kinds (all .npz, consumed by tests/ and by oracle pinning tests):
  tracks.npz             spline tables of the 5 shipped tracks as the reference builds them
                         (envs/tracks/tracks.py:14-64, torch_spline.py:240-301)
  lin_<case>.npz         DiffEnv.forward(cmd, approx='quad') dumps: traj, costs, transposed Jacobians,
                         cost gradients/Hessians, and the lazily evaluated dynamics-Hessian closures at
                         seeded random lambdas (envs/forward.py:188-321)
  lq_synth.npz           tests/test_auto_lin_quad.py: lin_quad_backward + roll_out_lin vs dense Newton
  steps_car.npz          tests/test_steps.py: the four directions on the simple car
  newton_pendulum.npz    tests/test_newton.py
  run_<case>_<algo>.npz  free-running run_min_algo traces with the iterate after every iteration
                         (teacher-forcing inputs) and the accepted internal line-search step
  mpc_<case>.npz         mpc_step windows (live) ; cached_*.npz excerpts of /root/reference/results
"""
import os
import sys
import types
import math
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def _stub():
    for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.collections", "matplotlib.patches", "seaborn"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib.collections"].PatchCollection = object
    sys.modules["matplotlib.collections"].LineCollection = object
    sys.modules["matplotlib.patches"].Polygon = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]


_stub()
sys.path.insert(0, REF)
import torch  # noqa: E402

torch.set_default_dtype(torch.float64)

from envs.pendulum import Pendulum, CartPendulum  # noqa: E402
from envs.car import Car  # noqa: E402
from envs.choose_env import make_env  # noqa: E402
from envs.backward import lin_quad_backward, quad_backward  # noqa: E402
from envs.rollout import roll_out_lin, roll_out_exact  # noqa: E402
from envs.lin_quad import make_synth_linear_env  # noqa: E402
import algorithms.algos_steps as algos_steps  # noqa: E402
import algorithms.run_min_algo as rma  # noqa: E402
from model_predictive_control.mpc_pipeline import mpc_step  # noqa: E402

rma.print_info_step = lambda *a, **k: None  # silence the table printer


def npf(x):
    if isinstance(x, torch.Tensor):
        return x.detach().numpy().astype(np.float64)
    return np.asarray(x, dtype=np.float64)


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrs)
    print("wrote", path, {k: getattr(v, "shape", None) for k, v in arrs.items()}, flush=True)


# ---------------------------------------------------------------------------------------------
# env cases: name -> (env_cfg for make_env, extra)
# ---------------------------------------------------------------------------------------------
ENV_CASES = {
    "pend_h100": dict(env="pendulum", horizon=100, dt=0.02),
    "pend_sp_h40": dict(env="pendulum", horizon=40, stay_put_time=1.0),
    "cart_h50": dict(env="cart_pendulum", horizon=50, dt=0.05, stay_put_time=0.6, x_limits=(-2.0, 2.0)),
    "cart_rk4cst_h25": dict(env="cart_pendulum", horizon=25, dt=0.1, stay_put_time=0.6, x_limits=(-2.0, 2.0),
                            discretization="rk4cst"),
    "cart_rk4_h10": dict(env="cart_pendulum", horizon=10, dt=0.1, stay_put_time=0.6, x_limits=(-2.0, 2.0),
                         discretization="rk4"),
    "scar_euler_exact_h50": dict(env="simple_car", track="simple", cost="exact", discretization="euler",
                                 reg_bar=0.0, horizon=50, dt=0.04),
    "scar_rk4cst_cont_h25": dict(env="simple_car", track="simple", cost="contouring",
                                 discretization="rk4_cst_ctrl", horizon=25, dt=0.04),
    "scar_rk4_exact_h10": dict(env="simple_car", track="simple", cost="exact", discretization="rk4",
                               reg_bar=0.0, horizon=10, dt=0.04),
    "bcar_h100": dict(env="real_car", track="simple", cost="contouring", discretization="rk4_cst_ctrl",
                      horizon=100, dt=0.01),
    "bcar_h25": dict(env="real_car", track="simple", cost="contouring", discretization="rk4_cst_ctrl",
                     horizon=25, dt=0.04),
    "bcar_euler_h20": dict(env="real_car", track="simple", cost="contouring", discretization="euler",
                           horizon=20, dt=0.02),
    "bcar_rk4_h10": dict(env="real_car", track="simple", cost="contouring", discretization="rk4",
                         horizon=10, dt=0.02),
    "bcar_complex_h40": dict(env="real_car", track="complex", vref=2.5, reg_cont=1.0, reg_lag=1.0,
                             reg_speed=1.0, horizon=40),
    "bcar_exact_h20": dict(env="real_car", track="simple", cost="exact", discretization="rk4_cst_ctrl",
                           reg_bar=0.0, horizon=20, dt=0.02),
    "bcar_circle_h20": dict(env="real_car", track="circle", horizon=20),
    # open tracks: handle_termination='repeat_end_point' -> smooth_min(t, L) (torch_spline.py:344-345)
    "bcar_line_h20": dict(env="real_car", track="line", horizon=20),
    "scar_bend_h20": dict(env="simple_car", track="bend", horizon=20, dt=0.04),
    # time_penalty='dref_squared' (car.py:213-216): constant (tref, vtref) cross term in the cost Hessian
    "bcar_dref2_h20": dict(env="real_car", track="simple", horizon=20, time_penalty="dref_squared"),
    "scar_tref_h20": dict(env="simple_car", track="simple", horizon=20, dt=0.04, time_penalty="tref"),
    "bcar_dref_h20": dict(env="real_car", track="circle", horizon=20, time_penalty="dref"),
}


def cfg_to_arrays(cfg):
    """Store the env_cfg as a repr string so tests can rebuild the same env through make_env."""
    return dict(env_cfg=np.array(repr(cfg)))


# ---------------------------------------------------------------------------------------------
def gen_tracks():
    from envs.tracks.tracks import get_track
    out = {}
    for name in ["simple", "complex", "circle", "line", "bend"]:
        center, inner, outer = get_track(name)
        out[name + "_knots"] = npf(center._t)
        for tag, sp in zip(["c", "i", "o"], [center, inner, outer]):
            assert torch.equal(sp._t, center._t)
            out[f"{name}_{tag}_a"] = npf(sp._a)
            out[f"{name}_{tag}_b"] = npf(sp._b)
            out[f"{name}_{tag}_c"] = npf(sp._c)
            out[f"{name}_{tag}_d"] = npf(sp._d)
        # probe values (evaluate / derivative incl. loop wrap and knot hits)
        L = float(max(center._t))
        ts = np.concatenate([np.linspace(-0.3 * L, 2.3 * L, 41), npf(center._t)[[0, 1, 5, -2, -1]]])
        ev = np.stack([npf(center.evaluate(torch.tensor(t))) for t in ts])
        de = np.stack([npf(center.derivative(torch.tensor(t))) for t in ts])
        evo = np.stack([npf(outer.evaluate(torch.tensor(t))) for t in ts])
        out[name + "_probe_t"] = ts
        out[name + "_probe_eval"] = ev
        out[name + "_probe_deriv"] = de
        out[name + "_probe_eval_outer"] = evo
    save("tracks", **out)


# ---------------------------------------------------------------------------------------------
def dump_forward(env, cmd, n_lam=2, seed=0):
    """forward(approx='quad') -> arrays (forward.py:100-139, :188-321)."""
    cmd = cmd.clone().requires_grad_(True)
    traj, costs = env.forward(cmd, approx="quad")
    H = cmd.shape[0]
    nx, nu = env.dim_state, env.dim_ctrl
    out = dict(
        cmd=npf(cmd), x0=npf(env.init_state), t0=np.array(env.init_time_iter),
        traj=np.stack([npf(s) for s in traj]),
        costs=np.array([float(c) for c in costs]),
        grad_dyn_state=np.stack([npf(traj[t + 1].grad_dyn_state) for t in range(H)]),
        grad_dyn_ctrl=np.stack([npf(traj[t + 1].grad_dyn_ctrl) for t in range(H)]),
        grad_state=np.stack([npf(c.grad_state) for c in costs]),
        hess_state=np.stack([npf(c.hess_state) for c in costs]),
        grad_ctrl=np.stack([npf(c.grad_ctrl).reshape(nu) for c in costs[1:]]),
        hess_ctrl=np.stack([npf(c.hess_ctrl).reshape(nu, nu) for c in costs[1:]]),
    )
    g = torch.Generator().manual_seed(1000 + seed)
    lams = torch.randn(n_lam, H, nx, generator=g)
    hxx = np.zeros((n_lam, H, nx, nx)); hxu = np.zeros((n_lam, H, nx, nu)); huu = np.zeros((n_lam, H, nu, nu))
    for i in range(n_lam):
        for t in range(H):
            if t > 0:  # traj[0] loses requires_grad in the reference (backward.py:102-108)
                hxx[i, t] = npf(traj[t + 1].hess_dyn_state(lams[i, t]))
            hxu[i, t] = npf(traj[t + 1].hess_dyn_state_ctrl(lams[i, t]))
            huu[i, t] = npf(traj[t + 1].hess_dyn_ctrl(lams[i, t]))
    out.update(lam=npf(lams), hess_dyn_state=hxx, hess_dyn_state_ctrl=hxu, hess_dyn_ctrl=huu)
    # backward passes on this expansion (backward.py:6-150) and both roll-outs (rollout.py)
    for reg in [0.0, 0.5]:
        tag = f"reg{reg:g}"
        gains, jcst, feas = lin_quad_backward(traj, costs, reg, handle_bad_dir="check_total_value")
        out[f"lq_{tag}_K"] = np.stack([npf(gn["gain_ctrl"]) for gn in gains])
        out[f"lq_{tag}_k"] = np.stack([npf(gn["offset_ctrl"]) for gn in gains])
        out[f"lq_{tag}_jcst"] = npf(jcst)
        out[f"lq_{tag}_rol"] = npf(roll_out_lin(traj, gains, 0.7))
        out[f"lq_{tag}_roe"] = npf(roll_out_exact(env, traj, gains, cmd, 0.7))
    for mode in ["newton", "ddp"]:
        for reg in [0.5, 20.0]:
            tag = f"{mode}_reg{reg:g}"
            gains, jcst, feas = quad_backward(traj, costs, reg, handle_bad_dir="check_total_value", mode=mode)
            out[f"q_{tag}_K"] = np.stack([npf(gn["gain_ctrl"]) for gn in gains])
            out[f"q_{tag}_k"] = np.stack([npf(gn["offset_ctrl"]) for gn in gains])
            out[f"q_{tag}_jcst"] = npf(jcst)
            out[f"q_{tag}_feasible"] = np.array(bool(feas))
            out[f"q_{tag}_rol"] = npf(roll_out_lin(traj, gains, 1.0))
            out[f"q_{tag}_roe"] = npf(roll_out_exact(env, traj, gains, cmd, 1.0))
    # gradient of the total cost wrt the controls (collect_info, run_min_algo.py:156-157)
    cmd2 = cmd.detach().clone().requires_grad_(True)
    _, c2 = env.forward(cmd2)
    out["grad_obj"] = npf(torch.autograd.grad(sum(c2), cmd2)[0])
    out["total_cost"] = npf(sum(c2)).reshape(())
    return out


def gen_lin(only=None):
    """Short-horizon expansions for every env/integrator/cost combination, generic states."""
    for name, cfg in ENV_CASES.items():
        if only and name not in only:
            continue
        cfg = dict(cfg)
        cfg["horizon"] = min(cfg["horizon"], 8)
        if name == "cart_h50":
            cfg["horizon"] = 14  # must straddle stay_put_time_start_iter = 14 - int(0.6/0.05) = 3
        env = make_env(cfg)
        g = torch.Generator().manual_seed(7)
        nx, nu = env.dim_state, env.dim_ctrl
        cmd = 0.5 * torch.randn(cfg["horizon"], nu, generator=g)
        # generic initial state so that no Jacobian entry is accidentally zero
        x0 = env.init_state.detach().clone()
        x0 = x0 + 0.1 * torch.randn(nx, generator=g)
        if cfg["env"] in ("real_car", "simple_car"):
            x0[-1] = abs(x0[-1]) + 0.5   # vtref>0 (log barrier)
            x0[-2] = 0.3                 # tref
            if cfg["env"] == "real_car":
                x0[3] = 1.2              # vx > 0
        if cfg.get("track") in ("line", "bend"):
            # start just before the end of the open track so that tref crosses it (and the 1e-6 blend zone
            # of smooth_min is left behind within the horizon)
            from envs.tracks.tracks import get_track
            x0[-2] = float(max(get_track(cfg["track"])[0]._t)) - 0.02
        if cfg["env"] == "cart_pendulum":
            x0[0] = 1.95                 # near the x_limits barrier so relu^3 is active on some steps
            cmd = cmd + 1.0
        env.init_state = x0
        t0 = time.time()
        out = dump_forward(env, cmd)
        out.update(cfg_to_arrays(cfg))
        save("lin_" + name, **out)
        print(name, "%.1fs" % (time.time() - t0), flush=True)
    if only:
        return
    # a car pushed outside the track so that both border barriers get exercised, and a window that
    # starts late (init_time_iter>0, exact cost depends on it, car.py:229)
    cfg = dict(ENV_CASES["bcar_h25"]); cfg["horizon"] = 6
    env = make_env(cfg)
    for tag, dy in [("in", 0.2), ("out", -0.2)]:
        x0 = env.init_state.detach().clone(); x0[1] = dy; x0[6] = 0.5
        env.init_state = x0
        g = torch.Generator().manual_seed(11)
        out = dump_forward(env, 0.3 * torch.randn(6, 3, generator=g))
        out.update(cfg_to_arrays(cfg))
        save("lin_bcar_border_" + tag, **out)
    cfg = dict(ENV_CASES["scar_euler_exact_h50"]); cfg["horizon"] = 6
    env = make_env(cfg); env.init_time_iter = 37
    g = torch.Generator().manual_seed(12)
    out = dump_forward(env, 0.3 * torch.randn(6, 3, generator=g))
    out.update(cfg_to_arrays(cfg))
    save("lin_scar_exact_t0_37", **out)


# ---------------------------------------------------------------------------------------------
def gen_lq():
    """tests/test_auto_lin_quad.py and tests/utils.py:direct_newton on the synthetic LQ env."""
    sys.path.insert(0, REF)
    from tests.utils import direct_newton
    torch.manual_seed(1)
    env = make_synth_linear_env(50, 4, 4)
    cmd0 = torch.rand(50, 4, requires_grad=True)
    traj, costs = env.forward(cmd0, approx="linquad")
    out = dict(
        A=npf(env.lin_dyn_states[0]), B=npf(env.lin_dyn_ctrls[0]),
        P=npf(env.quad_cost_states[1]), p=npf(env.lin_cost_states[1]),
        Q=npf(env.quad_cost_ctrls[0]), q=npf(env.lin_cost_ctrls[0]),
        x0=npf(env.init_state), cmd0=npf(cmd0), total_cost0=npf(sum(costs)),
        traj=np.stack([npf(s) for s in traj]),
        grad_state=np.stack([npf(c.grad_state) for c in costs]),
        grad_ctrl=np.stack([npf(c.grad_ctrl) for c in costs[1:]]),
    )
    for reg in [0.0, 10.0]:
        gains, jcst, feas = lin_quad_backward(traj, costs, reg)
        out[f"reg{reg:g}_K"] = np.stack([npf(g["gain_ctrl"]) for g in gains])
        out[f"reg{reg:g}_k"] = np.stack([npf(g["offset_ctrl"]) for g in gains])
        out[f"reg{reg:g}_jcst"] = npf(jcst)
        out[f"reg{reg:g}_diff"] = npf(roll_out_lin(traj, gains))
        dn, opt = direct_newton(env, cmd0, reg)
        out[f"reg{reg:g}_newton_diff"] = npf(dn)
        out[f"reg{reg:g}_newton_opt"] = npf(opt)
    save("lq_synth", **out)

    # tests/test_steps.py
    torch.manual_seed(0)
    env = Car(model="simple", discretization="euler", cost="exact", horizon=50, dt=0.02)
    ctrls = torch.randn(env.horizon, env.dim_ctrl, requires_grad=True)
    traj, costs = env.forward(ctrls, approx="linquad")
    pol = lin_quad_backward(traj, costs)[0]
    gn = roll_out_lin(traj, pol)
    ddp_lq = roll_out_exact(env, traj, pol, ctrls)
    traj, costs = env.forward(ctrls, approx="quad")
    pol = quad_backward(traj, costs, reg_ctrl=10.0)[0]
    nt = roll_out_lin(traj, pol)
    ddp_q = roll_out_exact(env, traj, pol, ctrls)
    save("steps_car", ctrls=npf(ctrls), gauss_newton_dir=npf(gn), ddp_linquad_dir=npf(ddp_lq),
         newton_dir=npf(nt), ddp_quad_dir=npf(ddp_q),
         norms=np.array([float(torch.norm(v)) for v in (gn, ddp_lq, nt, ddp_q)]))

    # tests/test_newton.py
    torch.manual_seed(1)
    env = Pendulum(horizon=100, stay_put_time=1.0)
    cmd = 0.1 * torch.randn(env.horizon, env.dim_ctrl, requires_grad=True)
    traj, costs = env.forward(cmd, approx="quad")
    o1 = algos_steps.classic_oracle(traj, costs, approx="quad", step_mode="dir", handle_bad_dir="check_total_value")
    o2 = algos_steps.newton_oracle(env, cmd)
    d1, v1 = o1(1.0)
    d2, v2 = o2(1.0)
    save("newton_pendulum", cmd=npf(cmd), ilqc_dir=npf(d1), ilqc_dec=npf(v1), autodiff_dir=npf(d2),
         autodiff_dec=npf(v2))


# ---------------------------------------------------------------------------------------------
class LSRecorder:
    """Records the internal accepted step of backtracking_line_search (algos_steps.py:384-437)."""

    def __init__(self):
        self.orig = algos_steps.backtracking_line_search
        self.accepted = []
        self.initial = []
        self.trips = []

    def __enter__(self):
        rec = self

        def wrapped(fun, var, oracle, stepsize, decrease_fac, slope_rtol=1.0):
            n = [0]

            def counting_oracle(s):
                n[0] += 1
                return oracle(s)

            rec.initial.append(float(stepsize))
            diff, s = rec.orig(fun, var, counting_oracle, stepsize, decrease_fac, slope_rtol)
            rec.accepted.append(float(s))
            rec.trips.append(n[0])
            return diff, s

        algos_steps.backtracking_line_search = wrapped
        return self

    def __exit__(self, *a):
        algos_steps.backtracking_line_search = self.orig


def trace_run(env, algo, n_iter, prev_cmd=None, stepsize=None, handle_bad_dir="check_total_value"):
    """Free-running run_min_algo, one iteration at a time through its own restart arguments
    (run_min_algo.py:50-75), recording the iterate after every iteration."""
    cmds, steps = [], []
    cmd = prev_cmd
    aux, metrics = None, None
    H, nu = env.horizon, env.dim_ctrl
    cmds.append(npf(cmd) if cmd is not None else np.zeros((H, nu)))
    with LSRecorder() as rec:
        for k in range(1, n_iter + 1):
            cmd, aux, metrics = rma.run_min_algo(
                env, max_iter=k, algo=algo, stepsize=stepsize if aux is None else None,
                handle_bad_dir=handle_bad_dir, prev_cmd=cmd, optim_aux_vars=aux, past_metrics=metrics)
            if cmd is None or metrics["iteration"][-1] < k:
                break
            cmds.append(npf(cmd))
            steps.append(float(aux["stepsize"]))
            cmd = cmd.detach().clone()
            if rma.check_cvg_status(metrics, verbose=False) != "running":
                break
    n = len(cmds) - 1
    return dict(
        cmds=np.stack(cmds), stepsize_state=np.array(steps),
        cost=np.array(metrics["cost"], dtype=np.float64),
        norm_grad_obj=np.array(metrics["norm_grad_obj"], dtype=np.float64),
        stepsize_reported=np.array([float(s) for s in metrics["stepsize"]], dtype=np.float64),
        ls_initial=np.array(rec.initial[:n]), ls_accepted=np.array(rec.accepted[:n]),
        ls_trips=np.array(rec.trips[:n]), algo=np.array(algo),
        status=np.array(rma.check_cvg_status(metrics, verbose=False)),
        handle_bad_dir=np.array(handle_bad_dir),
    )


RUNS = [
    # (case, algo, iters)
    ("pend_h100", "ddp_linquad_reg", 20),          # BASELINE config 1
    ("pend_h100", "classic_linquad_reg", 10),
    ("pend_h100", "ddp_linquad_dir", 10),
    ("pend_h100", "classic_linquad_dir", 10),
    ("pend_h100", "gd", 10),
    ("pend_h100", "classic_quad_reg", 8),
    ("pend_h100", "ddp_quad_reg", 8),
    ("pend_h100", "ddp_quad_dir", 6),
    ("pend_h100", "classic_quad_dir", 6),
    ("pend_h100", "ddp_linquad_regvar", 6),
    ("pend_h100", "ddp_linquad_dirvar", 6),
    ("cart_h50", "classic_linquad_reg", 12),       # BASELINE config 2
    ("cart_h50", "ddp_linquad_reg", 12),
    ("cart_h50", "ddp_quad_reg", 6),
    ("cart_h50", "gd", 6),
    ("cart_rk4cst_h25", "ddp_linquad_reg", 8),
    ("scar_euler_exact_h50", "ddp_linquad_reg", 10),  # BASELINE config 3
    ("scar_euler_exact_h50", "classic_linquad_reg", 6),
    ("scar_euler_exact_h50", "ddp_quad_reg", 4),
    ("scar_rk4cst_cont_h25", "ddp_linquad_reg", 6),
    ("bcar_h25", "ddp_linquad_reg", 6),
    ("bcar_h25", "classic_linquad_reg", 4),
    ("bcar_h25", "ddp_quad_reg", 3),
    ("bcar_h25", "classic_quad_reg", 3),
    ("bcar_h25", "ddp_linquad_dir", 3),
    ("bcar_euler_h20", "ddp_linquad_reg", 5),
    ("bcar_complex_h40", "ddp_linquad_reg", 4),
    ("bcar_h100", "ddp_linquad_reg", 3),           # BASELINE config 4 (headline algo)
    ("bcar_h100", "ddp_quad_reg", 1),              # BASELINE config 4 (nominal algo), ~9 min
    ("bcar_dref2_h20", "ddp_linquad_reg", 3),
]
# handle_bad_dir other than run_min_algo's default (backward.py:221-237): (case, algo, iters, handle_bad_dir)
RUNS_BAD_DIR = [
    ("pend_h100", "classic_quad_reg", 6, "flag"),
    ("pend_h100", "ddp_quad_reg", 6, "let_it_be"),
    ("cart_h50", "ddp_quad_reg", 5, "flag"),
    ("cart_h50", "ddp_linquad_reg", 5, "flag"),
    ("bcar_h25", "ddp_quad_reg", 2, "flag"),
]


def gen_runs(only=None):
    for case, algo, iters in RUNS:
        if only and f"{case}:{algo}" not in only and case not in only:
            continue
        env = make_env(dict(ENV_CASES[case]))
        t0 = time.time()
        out = trace_run(env, algo, iters)
        out.update(cfg_to_arrays(ENV_CASES[case]))
        out["x0"] = npf(env.init_state)
        save(f"run_{case}_{algo}", **out)
        print(case, algo, "%.1fs" % (time.time() - t0), flush=True)


def gen_runs_bad_dir():
    for case, algo, iters, hbd in RUNS_BAD_DIR:
        env = make_env(dict(ENV_CASES[case]))
        t0 = time.time()
        out = trace_run(env, algo, iters, handle_bad_dir=hbd)
        out.update(cfg_to_arrays(ENV_CASES[case]))
        out["x0"] = npf(env.init_state)
        save(f"run_{case}_{algo}__{hbd}", **out)
        print(case, algo, hbd, "%.1fs" % (time.time() - t0), out["ls_trips"], flush=True)


def gen_runs_perturbed():
    """Per-problem inputs as the batched workloads use them: x0 / vinit / weights."""
    g = torch.Generator().manual_seed(1234)
    xi = torch.randn(4, generator=g)
    for i in range(3):
        cfg = dict(ENV_CASES["cart_h50"])
        env = make_env(cfg)
        env.init_state = torch.tensor([float(xi[i]), 0.0, 0.0, 0.0])
        out = trace_run(env, "ddp_linquad_reg", 8)
        out.update(cfg_to_arrays(cfg)); out["x0"] = npf(env.init_state)
        save(f"run_cart_h50_x0_{i}_ddp_linquad_reg", **out)
    g = torch.Generator().manual_seed(1236)
    xi = torch.randn(2, 4, generator=g)
    for i in range(2):
        cfg = dict(ENV_CASES["bcar_h25"])
        cfg.update(vinit=1.0 + 0.01 * float(xi[i, 0]), reg_cont=0.1 + 1e-3 * float(xi[i, 1]),
                   reg_lag=10.0 + 1e-3 * float(xi[i, 2]), reg_speed=0.1 + 1e-3 * float(xi[i, 3]))
        env = make_env(cfg)
        out = trace_run(env, "ddp_linquad_reg", 4)
        out.update(cfg_to_arrays(cfg)); out["x0"] = npf(env.init_state)
        save(f"run_bcar_h25_pert_{i}_ddp_linquad_reg", **out)


# ---------------------------------------------------------------------------------------------
def gen_cached():
    """Excerpts of /root/reference/results that SURVEY.md section 4 found valid at HEAD."""
    def load(path):
        return torch.load(path, weights_only=False)

    d = os.path.join(REF, "results", "dt_2.00e-02_env_pendulum_horizon_100")
    cfg, outp = load(os.path.join(d, "algo_ddp_linquad_reg_max_iter_20.torch"))
    m = outp["metrics"]
    save("cached_pend_h100_ddp_linquad_reg_20", cmd_opt=npf(outp["cmd_opt"]), cost=np.array(m["cost"]),
         norm_grad_obj=np.array(m["norm_grad_obj"]),
         stepsize=np.array([float(s) for s in m["stepsize"]]))
    d = os.path.join(REF, "results",
                     "env_real_car_reg_cont_1.00e+00_reg_lag_1.00e+00_reg_speed_1.00e+00_track_complex_vref_2.50e+00")
    out = {}
    for h in [40, 41, 42, 43, 44, 100, 101, 400, 401]:
        cfg, outp = load(os.path.join(d, f"algo_ddp_linquad_reg_full_horizon_{h}_max_iter_10_overlap_39.torch"))
        out[f"cmd_opt_{h}"] = npf(outp["cmd_opt"])
        out[f"cost_{h}"] = np.array(outp["metrics"]["cost"])
    save("cached_mpc_complex", **out)


def gen_mpc():
    """Live mpc_step windows on the complex track (mpc_pipeline.py:18-55), chained like run_mpc."""
    cfg = dict(env="real_car", track="complex", vref=2.5, reg_cont=1.0, reg_lag=1.0, reg_speed=1.0)
    prev = None
    out = {}
    horizon = 0
    for w in range(4):
        horizon = max(horizon + 1, 40)
        env = make_env(cfg)
        t0 = time.time()
        cmd_opt, metrics = mpc_step(env, full_horizon=horizon, overlap=39, max_iter=10,
                                    algo="ddp_linquad_reg", prev_cmd=prev)
        out[f"cmd_opt_{w}"] = npf(cmd_opt)
        out[f"cost_{w}"] = np.array(metrics["cost"])
        out[f"full_horizon_{w}"] = np.array(horizon)
        out[f"x0_{w}"] = npf(env.init_state)
        out[f"t0_{w}"] = np.array(env.init_time_iter)
        prev = cmd_opt.detach().clone()
        print("mpc window", w, horizon, "%.1fs" % (time.time() - t0), metrics["cost"][-1], flush=True)
    out.update(cfg_to_arrays(cfg))
    save("mpc_complex_live", **out)


if __name__ == "__main__":
    groups = sys.argv[1:] or ["tracks", "lin", "lq", "runs", "pert", "cached", "mpc"]
    torch.set_num_threads(1)
    for gname in groups:
        if gname == "tracks":
            gen_tracks()
        elif gname == "lin":
            gen_lin()
        elif gname.startswith("lin:"):
            gen_lin(only=gname[4:].split(","))
        elif gname == "lq":
            gen_lq()
        elif gname == "baddir":
            gen_runs_bad_dir()
        elif gname.startswith("runs:"):
            gen_runs(only=gname[5:].split(","))
        elif gname == "runs":
            gen_runs()
        elif gname.startswith("runs:"):
            gen_runs(only=gname[5:].split(","))
        elif gname == "pert":
            gen_runs_perturbed()
        elif gname == "cached":
            gen_cached()
        elif gname == "mpc":
            gen_mpc()
        else:
            raise SystemExit("unknown group " + gname)
