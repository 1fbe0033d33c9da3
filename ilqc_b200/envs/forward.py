"""Host-side mirror of the reference's DiffEnv (envs/forward.py:11-186).

An env object only holds the parameters of one problem family (what the reference's constructors
take) plus `init_state` / `init_time_iter`; every computation is done by the CUDA kernels through
ilqc_b200.engine.  `forward(cmd, approx)` keeps the reference's return convention (lists `traj`,
`costs` whose items carry the expansion as attributes) so that code written against the reference -
tests, notebooks, the lower-level oracles - runs unchanged on top of the engine with a batch of one.
"""
import torch

from .. import _lib


class DiffEnv:
    """Base class.  Subclasses fill `_fill_desc` with their scalars (see pendulum.py, car.py)."""

    model_id = None

    def __init__(self):
        self.dim_ctrl, self.dim_state = None, None
        self.dt, self.horizon = None, None
        self.state, self.time_iter = None, 0
        self.init_state, self.init_time_iter = None, None
        self.viewer = None
        self.discretization = "euler"

    # ---- description handed to the C ABI ---------------------------------------------------------------
    _INTEG = {"euler": _lib.INT_EULER, "rk4_cst_ctrl": _lib.INT_RK4_CST, "rk4cst": _lib.INT_RK4_CST,
              "rk4": _lib.INT_RK4}

    def model_desc(self, horizon=None):
        """ctypes ModelDesc with the env scalars (device pointers are filled by the engine)."""
        d = _lib.ModelDesc()
        d.model = self.model_id
        if self.discretization not in self._INTEG:
            raise NotImplementedError(self.discretization)
        d.integrator = self._INTEG[self.discretization]
        d.horizon = int(self.horizon if horizon is None else horizon)
        d.t0 = int(self.init_time_iter)
        d.dt = float(self.dt)
        self._fill_desc(d)
        return d

    def _fill_desc(self, d):
        raise NotImplementedError

    track = None  # Car sets a ilqc_b200.envs.tracks.Track

    # ---- reference-style single-problem API (batch of one through the engine) --------------------------
    def _engine(self):
        from ..engine import default_engine
        return default_engine()

    def reset(self, requires_grad=False):
        self.time_iter = self.init_time_iter
        self.state = self.init_state
        return self.state

    def discrete_dyn(self, state, ctrl, time_iter):
        """One step of the dynamics x_{t+1} = phi(x_t, u_t) (envs/forward.py:29-43), on the device."""
        eng = self._engine()
        x0 = torch.as_tensor(state, dtype=torch.float64).reshape(1, -1)
        cmd = torch.as_tensor(ctrl, dtype=torch.float64).reshape(1, 1, -1)
        traj, _, _ = eng.rollout_cost(self, x0, cmd, t0=time_iter)
        return traj[0, 1].cpu()

    def step(self, ctrl, approx=None):
        traj, costs = self._forward_from(self.state, self.time_iter, torch.as_tensor(ctrl).reshape(1, -1), approx)
        self.state, self.time_iter = traj[1], self.time_iter + 1
        return traj[1], costs[1]

    def forward(self, cmd, approx=None, reset=True):
        """traj (H+1 states), costs (H+1 scalars, costs[0]=0); with approx in {'linquad','quad'} the items
        carry grad_dyn_state / grad_dyn_ctrl (transposed Jacobians), grad_state / hess_state, grad_ctrl /
        hess_ctrl, and for 'quad' the closures hess_dyn_state / hess_dyn_ctrl / hess_dyn_state_ctrl(lam)
        exactly as envs/forward.py:100-139, :188-321 store them."""
        if reset:
            self.reset()
        traj, costs = self._forward_from(self.state, self.time_iter, cmd, approx)
        self.state, self.time_iter = traj[-1], self.time_iter + cmd.shape[0]
        return traj, costs

    def _forward_from(self, state, time_iter, cmd, approx):
        eng = self._engine()
        cmd_t = torch.as_tensor(cmd, dtype=torch.float64).detach()
        H = cmd_t.shape[0]
        x0 = torch.as_tensor(state, dtype=torch.float64).detach().reshape(1, -1)
        if approx is None:
            tr, ct, _ = eng.rollout_cost(self, x0, cmd_t.reshape(1, H, -1), t0=time_iter)
            tr, ct = tr[0].cpu(), ct[0].cpu()
            traj, costs = [tr[t] for t in range(H + 1)], [ct[t] for t in range(H + 1)]
            traj[0]._ilqc = costs[0]._ilqc = dict(env=self, x0=x0, cmd=cmd_t, t0=time_iter, approx=None)
            return traj, costs
        if approx not in ("linquad", "quad"):
            raise NotImplementedError(approx)
        ex = eng.expand_dense(self, x0, cmd_t.reshape(1, H, -1), t0=time_iter)
        tr, ct, _ = eng.rollout_cost(self, x0, cmd_t.reshape(1, H, -1), t0=time_iter)
        tr, ct = tr[0].cpu(), ct[0].cpu()
        traj = [tr[t].clone() for t in range(H + 1)]
        costs = [ct[t].clone() for t in range(H + 1)]
        A, Bm = ex["A"][0].cpu(), ex["B"][0].cpu()
        for t in range(H):
            traj[t + 1].grad_dyn_state = A[t].t().contiguous()
            traj[t + 1].grad_dyn_ctrl = Bm[t].t().contiguous()
            costs[t + 1].grad_ctrl = ex["q"][0, t].cpu()
            costs[t + 1].hess_ctrl = ex["Q"][0, t].cpu()
        for t in range(H + 1):
            costs[t].grad_state = ex["p"][0, t].cpu()
            costs[t].hess_state = ex["P"][0, t].cpu()
        if approx == "quad":
            env = self
            for t in range(H):
                def _mk(t):
                    def hess(lam):
                        lam_full = torch.zeros(1, H, env.dim_state, dtype=torch.float64)
                        lam_full[0, t] = torch.as_tensor(lam, dtype=torch.float64)
                        e2 = eng.expand_dense(env, x0, cmd_t.reshape(1, H, -1), t0=time_iter, lam=lam_full)
                        return e2["Hxx"][0, t].cpu(), e2["Hxu"][0, t].cpu(), e2["Huu"][0, t].cpu()
                    return hess
                h = _mk(t)
                traj[t + 1].hess_dyn_state = (lambda lam, h=h: h(lam)[0])
                traj[t + 1].hess_dyn_state_ctrl = (lambda lam, h=h: h(lam)[1])
                traj[t + 1].hess_dyn_ctrl = (lambda lam, h=h: h(lam)[2])
        # what the batched kernels need to redo the expansion on the device
        traj[0]._ilqc = costs[0]._ilqc = dict(env=self, x0=x0, cmd=cmd_t, t0=time_iter, approx=approx)
        return traj, costs

    def close(self):
        self.viewer = None

    def visualize(self, cmd, dt=None):
        raise NotImplementedError("rendering is out of scope of the B200 engine (SURVEY.md section 2, #19)")

    def render(self, title=None):
        raise NotImplementedError("rendering is out of scope of the B200 engine (SURVEY.md section 2, #19)")
