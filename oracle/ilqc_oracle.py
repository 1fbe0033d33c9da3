"""ORACLE -- TEST INFRASTRUCTURE ONLY.  CPU float64 restatement of vroulet/ilqc's hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module, and only as the checker / the CPU baseline.  The product (ilqc_b200/) never imports it.

It restates, in plain PyTorch-CPU float64 with torch.autograd exactly where the reference uses it, the
algorithm of (file:line in the reference):
    DiffEnv.forward / step / diff_discrete_dyn / diff_cost      envs/forward.py:62-139, :188-321
    auto_multi_grad                                             envs/torch_utils.py:8-23
    euler / runge_kutta4_cst_ctrl                               envs/discretization.py:8-19, :42-57
    Pendulum, CartPendulum                                      envs/pendulum.py:56-87, :180-238
    Car.simple_model / bicycle_model / costs / barrier          envs/car.py:134-302
    NaturalCubicSpline.evaluate / derivative                    envs/tracks/torch_spline.py:341-375
    lin_quad_backward / quad_backward / bell_step / bell_core   envs/backward.py:6-257
    roll_out_lin / roll_out_exact                               envs/rollout.py:8-88
    gd_oracle / classic_oracle / ddp_oracle / min_step          algorithms/algos_steps.py:22-43, :105-381
    backtracking_line_search                                    algorithms/algos_steps.py:384-437
    run_min_algo / collect_info / check_cvg_status              algorithms/run_min_algo.py:15-264
    mpc_step                                                    model_predictive_control/mpc_pipeline.py:18-55

PINNED: tests/test_oracle_pinned.py checks this module against golden vectors produced by running the
unmodified reference in the build container (tests/golden/make_golden.py; fixtures committed), including
the reference's own known-answer tests (tests/test_auto_lin_quad.py, test_newton.py, test_steps.py), the
cached result file of BASELINE config 1 and the cached complex-track MPC windows.
"""
import math
import os
import sys
import time as _time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# the track way-points / spline fit and the env parameter containers are shared host-side data code
from ilqc_b200.envs.tracks import get_track  # noqa: E402
from ilqc_b200.envs.car import BICYCLE_CONSTANTS  # noqa: E402

torch.set_default_dtype(torch.float64)
F64 = torch.float64


def T(x):
    return torch.as_tensor(x, dtype=F64)


# ------------------------------------------------------------------------------------------------------
# autograd helper (torch_utils.py:8-23): transposed Jacobian, one reverse sweep per output row
# ------------------------------------------------------------------------------------------------------
def jac_t(output, inp, create_graph=False):
    rows = []
    for e in torch.eye(len(output)):
        rows.append(torch.autograd.grad(output, inp, e, retain_graph=True, create_graph=create_graph)[0])
    return torch.stack(rows).t()


# ------------------------------------------------------------------------------------------------------
# splines (torch_spline.py:341-375)
# ------------------------------------------------------------------------------------------------------
def smooth_relu(x, eps=1e-6):
    """torch_utils.py:44-69: cubic blend between 0 and x on [-eps, eps], branches selected by boolean masks."""
    x1, y1, x2, y2 = -eps, T(0.0), eps, eps
    k1, k2 = T(0.0), T(1.0)
    a = k1 * (x2 - x1) - (y2 - y1)
    b = -k2 * (x2 - x1) + (y2 - y1)
    tx = (x - x1) / (x2 - x1)
    return (x > eps) * x + (x >= -eps) * (x <= eps) * ((1 - tx) * y1 + tx * y2 + tx * (1 - tx) * ((1 - tx) * a + tx * b))


def smooth_min(x, a):
    return -smooth_relu(-x + a) + a


class Spline:
    def __init__(self, track, which):
        self.t = T(track.knots)
        c = T(track.coef[:, which])  # [n_seg, 2, 4]
        self.a, self.b, self.c, self.d = c[:, :, 0], c[:, :, 1], c[:, :, 2], c[:, :, 3]
        self.loop = track.loop
        self.length = float(track.knots[-1])

    def _locate(self, t):
        if self.loop:
            t = t % self.length
        else:
            t = smooth_min(t, self.length)  # 'repeat_end_point' (torch_spline.py:344-345)
        idx = torch.bucketize(t.detach(), self.t) - 1
        idx = idx.clamp(0, self.b.size(0) - 1)
        return t - self.t[idx], idx

    def evaluate(self, t):
        f, i = self._locate(t)
        inner = self.c[i] + self.d[i] * f
        inner = self.b[i] + inner * f
        return self.a[i] + inner * f

    def derivative(self, t):
        f, i = self._locate(t)
        inner = 2 * self.c[i] + 3 * self.d[i] * f
        return self.b[i] + inner * f


# ------------------------------------------------------------------------------------------------------
# integrators (discretization.py)
# ------------------------------------------------------------------------------------------------------
def euler(dyn, x, u, dt):
    return x + dt * dyn(x, u)


def rk4_cst(dyn, x, u, dt):
    k1 = dyn(x, u)
    k2 = dyn(x + dt * k1 / 2, u)
    k3 = dyn(x + dt * k2 / 2, u)
    k4 = dyn(x + dt * k3, u)
    return x + dt * (k1 + 2 * k2 + 2 * k3 + k4) / 6


def rk4_var(dyn, x, u, dt):
    """runge_kutta4 with three control sets per step (discretization.py:22-39)."""
    n = int(u.shape[0] / 3)
    k1 = dyn(x, u[:n])
    k2 = dyn(x + dt * k1 / 2, u[n:2 * n])
    k3 = dyn(x + dt * k2 / 2, u[n:2 * n])
    k4 = dyn(x + dt * k3, u[2 * n:])
    return x + dt * (k1 + 2 * k2 + 2 * k3 + k4) / 6


INTEGRATORS = {"euler": euler, "rk4_cst_ctrl": rk4_cst, "rk4cst": rk4_cst, "rk4": rk4_var}


# ------------------------------------------------------------------------------------------------------
# environments
# ------------------------------------------------------------------------------------------------------
class Env:
    def __init__(self):
        self.init_time_iter = 0

    def discrete_dyn(self, x, u, t):
        return INTEGRATORS[self.discretization](self.dyn, x, u, self.dt)


class Pendulum(Env):
    def __init__(self, reg_speed=0.1, reg_ctrl=1e-6, dt=0.05, horizon=40, stay_put_time=None, seed=0):
        super().__init__()
        self.horizon, self.dt, self.dim_ctrl, self.dim_state = horizon, dt, 1, 2
        self.discretization = "euler"
        self.init_state = T([math.pi, 0.0])
        if seed != 0:
            torch.manual_seed(seed)
            self.init_state[0] = self.init_state[0] + torch.randn(1)
        self.reg_speed, self.reg_ctrl = reg_speed, reg_ctrl
        self.start = horizon - int(stay_put_time / dt) if stay_put_time is not None else horizon
        self.g, self.m, self.l, self.mu = 10, 1.0, 1.0, 0.01

    def dyn(self, x, u):
        th, thdot = x
        g, m, l, mu = self.g, self.m, self.l, self.mu
        dthdot = -g / l * torch.sin(th + math.pi) - mu / (m * l ** 2) * thdot + 1 / (m * l ** 2) * u
        return torch.stack((thdot.unsqueeze(-1), dthdot)).view(-1)

    def costs(self, xn, u, t):
        cs = xn[0] ** 2 + self.reg_speed * xn[1] ** 2 if t >= self.start else T(0.0)
        return cs, self.reg_ctrl * u ** 2


class CartPendulum(Env):
    def __init__(self, reg_speed=0.1, reg_ctrl=1e-6, reg_barrier=1.0, dt=0.05, horizon=40, stay_put_time=None,
                 discretization="euler", x_limits=None, seed=0):
        super().__init__()
        self.horizon, self.dt, self.discretization = horizon, dt, discretization
        self.dim_ctrl, self.dim_state = (3 if discretization == "rk4" else 1), 4
        self.init_state = torch.zeros(4)
        if seed != 0:
            torch.manual_seed(seed)
            self.init_state[0] = self.init_state[0] + torch.randn(1)
        self.reg_speed, self.reg_ctrl, self.reg_barrier = reg_speed, reg_ctrl, reg_barrier
        self.start = horizon - int(stay_put_time / dt) if stay_put_time is not None else horizon
        self.g, self.M, self.m, self.b, self.I, self.l = 10, 0.5, 0.2, 0.1, 0.006, 0.3
        self.x_limits = x_limits

    def dyn(self, s, u):
        x, xdot, th, thdot = s  # literal unpacking order of the reference (pendulum.py:181)
        g, M, m, b, I, l = self.g, self.M, self.m, self.b, self.I, self.l
        a11, a22 = M + m, I + m * l ** 2
        a12 = m * l * torch.cos(th)
        detA = I * (M + m) + m * l ** 2 * M + m ** 2 * l ** 2 * torch.sin(th) ** 2
        b1 = m * l * thdot ** 2 * torch.sin(th) - b * xdot + u
        b2 = -m * g * l * torch.sin(th)
        dxdot = (a22 * b1 - a12 * b2) / detA
        dthdot = (-a12 * b1 + a11 * b2) / detA
        return torch.stack((xdot.unsqueeze(-1), thdot.unsqueeze(-1), dxdot, dthdot)).view(-1)

    def costs(self, xn, u, t):
        cs = (xn[1] + math.pi) ** 2 + self.reg_speed * xn[3] ** 2 if t >= self.start else T(0.0)
        if self.x_limits is not None:
            cs = cs + self.reg_barrier * (torch.relu(-(xn[0] - self.x_limits[0])) ** 3
                                          + torch.relu(-(self.x_limits[1] - xn[0])) ** 3)
        return cs, self.reg_ctrl * u.dot(u)


class Car(Env):
    def __init__(self, dt=0.02, discretization="rk4_cst_ctrl", horizon=20, track="simple", model="simple",
                 reg_ctrl=1e-6, reg_cont=0.1, reg_lag=10.0, reg_speed=0.1, reg_bar=100.0, reg_obs=0.0,
                 cost="contouring", time_penalty="vref_squared", vref=3.0, vinit=1.0,
                 constrain_angle=math.pi / 3, acc_bounds=(-0.1, 1.0), seed=0):
        super().__init__()
        if seed != 0:
            torch.manual_seed(seed)
            vinit = vinit + 1e-2 * torch.randn(1).item()
            reg_cont, reg_lag, reg_speed = (reg_cont + 1e-3 * torch.randn(1).item(),
                                            reg_lag + 1e-3 * torch.randn(1).item(),
                                            reg_speed + 1e-3 * torch.randn(1).item())
        self.dt, self.discretization, self.horizon, self.model = dt, discretization, horizon, model
        self.vref, self.constrain_angle, self.acc_bounds = vref, constrain_angle, acc_bounds
        self.reg_ctrl, self.reg_cont, self.reg_lag = reg_ctrl, reg_cont, reg_lag
        self.reg_speed, self.reg_bar, self.reg_obs = reg_speed, reg_bar, reg_obs
        self.cost_type, self.time_penalty = cost, time_penalty
        tr = get_track(track)
        self.track, self.inner, self.outer = Spline(tr, 0), Spline(tr, 1), Spline(tr, 2)
        dx_ref, dy_ref = self.track.derivative(T(0.0))
        phi_ref = torch.atan2(dy_ref, dx_ref)
        if model == "simple":
            self.init_state = T([0.0, 0.0, phi_ref, vinit, 0.0, vinit])
            self.dim_ctrl, self.dim_state = 3, 6
            self.dyn = self.simple_model
        else:
            self.init_state = T([0.0, 0.0, phi_ref, vinit, 0.0, 0.0, 0.0, vinit])
            self.dim_ctrl, self.dim_state = 3, 8
            self.dyn = self.bicycle_model
        if discretization == "rk4":
            self.dim_ctrl *= 3
        self.c = BICYCLE_CONSTANTS
        self.obstacle = self.track.evaluate(T(tr.length / 20)) + T([0.2, 0.0])

    def simple_model(self, s, u):
        x, y, phi, v, tref, vtref = s
        a, delta, atref = u
        delta = 2 / math.pi * torch.arctan(delta) * self.constrain_angle
        carlength = 1.5 * self.c["car_l"]
        return torch.stack([v * torch.cos(phi), v * torch.sin(phi), v * torch.tan(delta) / carlength, a, vtref, atref])

    def bicycle_model(self, s, u):
        x, y, phi, vx, vy, vphi, tref, vtref = s
        a, delta, atref = u
        delta = 2 / math.pi * torch.arctan(delta) * self.constrain_angle
        gap = self.acc_bounds[1] - self.acc_bounds[0]
        a = gap * torch.sigmoid(4 * a / gap) + self.acc_bounds[0]
        c = self.c
        alphaf = delta - torch.atan2(vphi * c["lf"] + vy, vx)
        alphar = torch.atan2(vphi * c["lr"] - vy, vx)
        Fry = c["Dr"] * torch.sin(c["Cr"] * torch.atan(c["Br"] * alphar))
        Ffy = c["Df"] * torch.sin(c["Cf"] * torch.atan(c["Bf"] * alphaf))
        Frx = (c["Cm1"] - c["Cm2"] * vx) * a - c["Cr0"] - c["Cr2"] * vx ** 2
        m = c["m"]
        dx = torch.cos(phi) * vx - torch.sin(phi) * vy
        dy = torch.sin(phi) * vx + torch.cos(phi) * vy
        dvx = (Frx - Ffy * torch.sin(delta) + m * vy * vphi) / m
        dvy = (Fry + Ffy * torch.cos(delta) - m * vx * vphi) / m
        dvphi = (Ffy * c["lf"] * torch.cos(delta) - Fry * c["lr"]) / c["Iz"]
        return torch.stack([dx, dy, vphi, dvx, dvy, dvphi, vtref, atref])

    def costs(self, xn, u, t):
        x, y = xn[0], xn[1]
        tref, vtref = xn[-2], xn[-1]
        if self.cost_type == "contouring":
            x_ref, y_ref = self.track.evaluate(tref)
            dx_ref, dy_ref = self.track.derivative(tref)
            phi_ref = torch.atan2(dy_ref, dx_ref)
            cont = torch.sin(phi_ref) * (x - x_ref) - torch.cos(phi_ref) * (y - y_ref)
            lag = -torch.cos(phi_ref) * (x - x_ref) - torch.sin(phi_ref) * (y - y_ref)
            cs = self.reg_cont * cont ** 2 + self.reg_lag * lag ** 2
            if self.time_penalty == "tref":
                cs = cs - self.reg_speed * tref
            elif self.time_penalty == "dref":
                cs = cs - self.reg_speed * vtref * self.dt
            elif self.time_penalty == "dref_squared":
                cs = cs - self.reg_speed * vtref * self.dt * tref
            elif self.time_penalty == "vref_squared":
                cs = cs + self.reg_speed * (vtref - self.vref) ** 2 * self.dt ** 2
            else:
                raise NotImplementedError
            cs = cs + self.barrier(xn)
        else:
            x_ref, y_ref = self.track.evaluate(T(t * self.vref * self.dt))
            cs = (x - x_ref) ** 2 + (y - y_ref) ** 2
        if isinstance(self.reg_ctrl, tuple):
            cc = self.reg_ctrl[0] * ((u[::3] ** 2).sum() + (u[1::3] ** 2).sum()) + self.reg_ctrl[1] * (u[2::3] ** 2).sum()
        else:
            cc = self.reg_ctrl * u.dot(u)
        return cs, cc

    def barrier(self, xn):
        bar = -1e-6 * torch.log(xn[-1])
        width = 1.5 * self.c["car_w"]
        tref = xn[-2]
        point = torch.stack([xn[0], xn[1]]).detach()  # torch.tensor([x, y]) in the reference: no gradient
        for i, border in enumerate([self.inner, self.outer]):
            bp = border.evaluate(tref)
            dpos = border.derivative(tref)
            v = torch.sqrt(torch.sum(dpos ** 2))
            normal = torch.stack([dpos[1], -dpos[0]]).detach() / v  # numerator detached in the reference
            if i == 0:
                bar = bar + self.reg_bar * torch.maximum(width / 2 - (point - bp).dot(normal), T(0.0)) ** 3
            else:
                bar = bar + self.reg_bar * torch.maximum((point - bp).dot(normal) + width / 2, T(0.0)) ** 3
        if self.reg_obs > 0.0:
            bar = bar + self.reg_obs * torch.maximum(width ** 2 - torch.sum((point - self.obstacle) ** 2), T(0.0)) ** 3
        return bar


def make_env(cfg):
    opts = {k: v for k, v in cfg.items() if k != "env"}
    kind = cfg["env"]
    if kind == "pendulum":
        return Pendulum(**opts)
    if kind == "cart_pendulum":
        return CartPendulum(**opts)
    if kind == "real_car":
        return Car(model="real", **opts)
    if kind == "simple_car":
        return Car(model="simple", **opts)
    raise NotImplementedError(kind)


# ------------------------------------------------------------------------------------------------------
# forward pass with recorded expansions (forward.py)
# ------------------------------------------------------------------------------------------------------
class Step:
    """what the reference stores as attributes of traj[t+1] / costs[t+1]"""
    __slots__ = ("A", "B", "p", "P", "q", "Q", "hxx", "huu", "hxu", "graph")


def forward(env, cmd, approx=None):
    """-> traj (list of H+1 tensors), costs (list of H+1 0-d/1-d tensors), steps (list of H Step or None)."""
    nx, nu, H = env.dim_state, env.dim_ctrl, cmd.shape[0]
    x = env.init_state.detach().clone()
    if approx is not None:
        x.requires_grad_(True)
        if not cmd.requires_grad:
            cmd = cmd.detach().clone().requires_grad_(True)
    traj, costs, steps = [x], [T(0.0)], []
    t = env.init_time_iter
    for i in range(H):
        u = cmd[i]
        if approx is None:
            xn = env.discrete_dyn(x, u, t)
            cs, cc = env.costs(xn, u, t + 1)
            costs.append(sum((cs, cc)))
        else:
            quad = approx == "quad"
            xn = env.discrete_dyn(x, u, t)
            st = Step()
            gx = jac_t(xn, x, create_graph=quad)
            gu = jac_t(xn, u, create_graph=quad)
            st.A, st.B = gx.detach().t(), gu.detach().t()
            if quad:
                st.graph = (gx, gu, x, u)
            cs, cc = env.costs(xn, u, t + 1)
            cost = cs + cc
            if cs.grad_fn is not None:
                g = torch.autograd.grad(cs, xn, create_graph=True)[0]
                st.P, st.p = jac_t(g, xn).detach(), g.detach()
            else:
                st.P, st.p = torch.zeros(nx, nx), torch.zeros(nx)
            if cc.grad_fn is not None:
                g = torch.autograd.grad(cc, u, create_graph=True)[0]
                st.Q, st.q = jac_t(g, u).detach(), g.detach()
            else:
                st.Q, st.q = torch.zeros(nu, nu), torch.zeros(nu)
            steps.append(st)
            costs.append(cost)
        x = xn
        t += 1
        traj.append(x)
    return traj, costs, steps


def hess_dyn(st, lam, nx, nu):
    """the three closures of forward.py:218-269 evaluated at lam: (Hxx, Huu, Hxu)."""
    gx, gu, x, u = st.graph

    def safe(fn, shape):
        try:
            return fn()
        except RuntimeError as err:
            if "One of the differentiated Tensors" in str(err) or "does not require grad" in str(err):
                return torch.zeros(*shape)
            raise

    if gx.grad_fn is None:
        hxx, hxu = torch.zeros(nx, nx), torch.zeros(nx, nu)
    else:
        hxx = safe(lambda: jac_t(gx.mv(lam), x), (nx, nx))
        hxu = safe(lambda: jac_t(gx.mv(lam), u).t(), (nx, nu))
    huu = torch.zeros(nu, nu) if gu.grad_fn is None else safe(lambda: jac_t(gu.mv(lam), u), (nu, nu))
    return hxx.detach(), huu.detach(), hxu.detach()


# ------------------------------------------------------------------------------------------------------
# backward passes (backward.py)
# ------------------------------------------------------------------------------------------------------
def bell(A, B, J, j, P, p, Q, q, R=None):
    Hm = Q + B.t().mm(J.mm(B))
    if R is None:
        R = torch.zeros(p.shape[0], q.shape[0])
    K = -torch.linalg.solve(Hm, R.t() + B.t().mm(J.mm(A)))
    k = -torch.linalg.solve(Hm, q + B.t().mv(j))
    Jn = P + A.t().mm(J.mm(A)) + (R + A.t().mm(J.mm(B))).mm(K)
    jn = p + A.t().mv(j) + A.t().mv(J.mv(B.mv(k))) + R.mv(k)
    rest = 0.5 * (q + B.t().mv(j)).dot(k)
    return Jn, jn, rest, K, k


def backward(steps, reg_ctrl=0.0, reg_state=0.0, mode="linquad", handle_bad_dir="check_total_value"):
    """lin_quad_backward (mode='linquad') / quad_backward (mode in 'newton','ddp') -> gains, jcst, feasible."""
    H = len(steps)
    nx, nu = steps[0].A.shape[0], steps[0].B.shape[1]
    J, j, jcst = steps[-1].P + reg_state * torch.eye(nx), steps[-1].p, 0.0
    lam = j
    gains = [None] * H
    feasible = True
    for t in range(H - 1, -1, -1):
        st = steps[t]
        P, p = (steps[t - 1].P, steps[t - 1].p) if t > 0 else (torch.zeros(nx, nx), torch.zeros(nx))
        Q, q = st.Q, st.q
        R = None
        if mode == "linquad":
            P = P + reg_state * torch.eye(nx)
            Q = Q + reg_ctrl * torch.eye(nu)
        else:
            hxx, huu, hxu = hess_dyn(st, lam, nx, nu)
            if t > 0:
                P = P + hxx + reg_state * torch.eye(nx)
            R = hxu
            Q = Q + huu + reg_ctrl * torch.eye(nu)
        J, jn, rest, K, k = bell(st.A, st.B, J, j, P, p, Q, q, R)
        if rest > 0 and handle_bad_dir == "flag":
            return gains, None, False
        jcst = jcst + rest
        gains[t] = (K, k)
        if mode == "newton":
            lam = p + st.A.t().mv(lam)
        j = jn
        if mode == "ddp":
            lam = j
    if mode != "linquad" and handle_bad_dir == "check_total_value":
        feasible = bool(jcst < 0)
    return gains, jcst, feasible


# ------------------------------------------------------------------------------------------------------
# roll-outs (rollout.py)
# ------------------------------------------------------------------------------------------------------
def roll_out_lin(steps, gains, stepsize=1.0):
    nx = steps[0].A.shape[0]
    dx = torch.zeros(nx)
    out = []
    for st, (K, k) in zip(steps, gains):
        v = K.mv(dx) + stepsize * k
        out.append(v)
        dx = st.A.mv(dx) + st.B.mv(v)
    return torch.stack(out)


def roll_out_exact(env, traj, gains, cmd, stepsize=1.0):
    dx, x = torch.zeros(env.dim_state), env.init_state.detach()
    out = []
    with torch.no_grad():
        for t, (K, k) in enumerate(gains):
            v = K.mv(dx) + stepsize * k
            out.append(v)
            x = env.discrete_dyn(x, cmd[t].detach() + v, t)
            dx = x - traj[t + 1].detach()
    return torch.stack(out)


# ------------------------------------------------------------------------------------------------------
# one optimisation step (algos_steps.py)
# ------------------------------------------------------------------------------------------------------
def objective(env, cmd):
    with torch.no_grad():
        _, costs, _ = forward(env, cmd)
    return sum(costs)


def grad_objective(env, cmd):
    c = cmd.detach().clone().requires_grad_(True)
    _, costs, _ = forward(env, c)
    return torch.autograd.grad(sum(costs), c)[0]


def parse_algo(algo):
    if algo == "gd":
        return "classic", None, "dir"
    algo_type, approx, step_mode = algo.split("_")
    assert algo_type in ("classic", "ddp") and approx in ("linquad", "quad")
    assert step_mode in ("dir", "reg", "dirvar", "regvar")
    return algo_type, approx, step_mode


def make_oracle(env, traj, costs, steps, cmd, algo, handle_bad_dir):
    if algo == "gd":
        g = grad_objective(env, cmd)
        fixed = -0.5 * torch.sum(g * g)
        return lambda s: (-s * g, s * fixed)
    algo_type, approx, step_mode = parse_algo(algo)
    mode = "linquad" if approx == "linquad" else ("newton" if algo_type == "classic" else "ddp")
    if step_mode in ("dir", "dirvar"):
        reg = 0.0
        gains, opt, feas = backward(steps, reg, mode=mode, handle_bad_dir=handle_bad_dir)
        while not feas and reg < 1e6:
            reg = 10.0 * reg if reg > 0.0 else 1e-6
            gains, opt, feas = backward(steps, reg, mode=mode, handle_bad_dir=handle_bad_dir)
        if algo_type == "classic":
            d = roll_out_lin(steps, gains) if feas else None
            return lambda s: (s * d, s * opt) if feas else (None, None)
        return lambda s: (roll_out_exact(env, traj, gains, cmd, s), s * opt) if feas else (None, None)

    def oracle(s):
        gains, opt, feas = backward(steps, 1 / s, mode=mode, handle_bad_dir=handle_bad_dir)
        if not feas:
            return None, None
        if algo_type == "classic":
            return roll_out_lin(steps, gains), opt
        return roll_out_exact(env, traj, gains, cmd, 1.0), opt
    return oracle


def line_search(fun, var, oracle, stepsize, rho, trips=None):
    direction = float(np.sign(stepsize))
    curr = fun(var)
    n = 0
    while True:
        diff, dec = oracle(stepsize)
        n += 1
        if diff is None:
            decrease, stop = False, abs(stepsize) < 1e-32
        else:
            decrease = bool(direction * (fun(var + diff) - (curr + dec)) < 0.0)
            stop = bool(torch.norm(diff) < 1e-32)
        if decrease or stop:
            break
        stepsize = stepsize * rho
    if trips is not None:
        trips.append(n)
    return diff, stepsize


def min_step(env, traj, costs, steps, cmd, stepsize, algo, line_search_on=True,
             handle_bad_dir="check_total_value", trips=None):
    oracle = make_oracle(env, traj, costs, steps, cmd, algo, handle_bad_dir)
    step_mode = "regvar" if algo == "gd" else parse_algo(algo)[2]
    approx = None if algo == "gd" else parse_algo(algo)[1]
    if line_search_on:
        if step_mode == "dir":
            stepsize, rho = 1.0, 0.9
        elif step_mode == "reg":
            ng = torch.sqrt(sum([st.q.reshape(-1).dot(st.q.reshape(-1)) + st.p.dot(st.p) for st in steps]))
            stepsize, rho = 8 * stepsize / ng, 0.5
        elif step_mode == "regvar":
            stepsize, rho = stepsize * 8, 0.5
        else:
            stepsize, rho = min(4 * stepsize, 1), 0.9
        diff, stepsize = line_search(lambda c: objective(env, c), cmd.detach(), oracle, stepsize, rho, trips)
        if step_mode == "reg":
            stepsize = stepsize * ng
    else:
        diff, _ = oracle(stepsize)
    if diff is None:
        return None, None, None, None, stepsize
    nxt = (cmd.detach() + diff).requires_grad_(True)
    traj, costs, steps = forward(env, nxt, approx=approx)
    return nxt, traj, costs, steps, stepsize


# ------------------------------------------------------------------------------------------------------
# solver loop (run_min_algo.py)
# ------------------------------------------------------------------------------------------------------
def check_status(metrics):
    c = metrics["cost"]
    status = "running"
    if math.isnan(c[-1]) or c[-1] > c[0] + 1e3:
        status = "diverged"
    if len(c) > 1 and abs(c[-1] - c[-2]) / abs(c[-2]) < 1e-24:
        status = "converged"
    if status == "running" and metrics["stepsize"][-1] < 1e-24:
        status = "got stuck"
    return status


def run_min_algo(env, max_iter=10, algo="ddp_linquad_reg", stepsize=None, line_search=True,
                 handle_bad_dir="check_total_value", prev_cmd=None, optim_aux_vars=None, past_metrics=None,
                 verbose=False, record=None):
    cmd = prev_cmd.detach().clone() if prev_cmd is not None else torch.zeros(env.horizon, env.dim_ctrl)
    cmd.requires_grad_(True)
    step_mode = "dir" if algo == "gd" else parse_algo(algo)[2]
    approx = None if algo == "gd" else parse_algo(algo)[1]
    if optim_aux_vars is not None:
        stepsize = optim_aux_vars["stepsize"]
    elif stepsize is None:
        stepsize = 100.0 if "reg" in step_mode else 1.0
    traj, costs, steps = forward(env, cmd, approx=approx)

    def collect(metrics, it, cmd, costs, stepsize, dt):
        cost = float(sum(costs)) if costs is not None else float("nan")
        gn = float(torch.norm(grad_objective(env, cmd)))
        rep = stepsize
        if algo != "gd" and step_mode == "reg" and it > 0:
            rep = float(stepsize * metrics["norm_grad_obj"][-1])
        row = dict(iteration=it, time=(metrics["time"][-1] if metrics else 0.0) + dt, cost=cost, norm_grad_obj=gn,
                   stepsize=float(rep), algo=algo)
        if metrics is None:
            return {k: [v] for k, v in row.items()}
        for k, v in row.items():
            metrics[k].append(v)
        return metrics

    if past_metrics is None:
        metrics, it = collect(None, 0, cmd, costs, stepsize, 0.0), 0
    else:
        metrics, it = past_metrics, past_metrics["iteration"][-1]
    trips = []
    while check_status(metrics) == "running" and it < max_iter:
        t0 = _time.time()
        nxt, traj_n, costs_n, steps_n, stepsize = min_step(env, traj, costs, steps, cmd, stepsize, algo, line_search,
                                                         handle_bad_dir, trips)
        it += 1
        if nxt is None:
            metrics["cost"].append(float("nan"))
            break
        cmd, traj, costs, steps = nxt, traj_n, costs_n, steps_n
        metrics = collect(metrics, it, cmd, costs, stepsize, _time.time() - t0)
        if record is not None:
            record.append(cmd.detach().clone())
        if verbose:
            print(it, metrics["cost"][-1], metrics["stepsize"][-1])
    metrics["ls_trips"] = trips
    return cmd.detach(), dict(stepsize=stepsize), metrics


# ------------------------------------------------------------------------------------------------------
# MPC (mpc_pipeline.py:18-55)
# ------------------------------------------------------------------------------------------------------
def mpc_step(env, full_horizon, overlap, max_iter=10, algo="ddp_linquad_reg", prev_cmd=None):
    curr = len(prev_cmd) if prev_cmd is not None else 0
    extra = full_horizon - curr
    if prev_cmd is None:
        cmd, _, metrics = run_min_algo(env, max_iter, algo)
        return cmd, metrics
    init_cmd = torch.cat((prev_cmd[curr - overlap:], prev_cmd[-1].repeat(extra, 1)))
    with torch.no_grad():
        traj, _, _ = forward(env, prev_cmd)
    env.init_state = traj[curr - overlap].detach()
    env.init_time_iter = curr - overlap
    env.horizon = init_cmd.shape[0]
    cmd, _, metrics = run_min_algo(env, max_iter, algo, prev_cmd=init_cmd)
    return torch.cat((prev_cmd[:-overlap], cmd)), metrics
