"""Thin host layer over the C ABI: torch tensors in, torch tensors out.

User-facing tensors are batch-major like the reference's per-problem tensors with a leading batch axis
(`x0 [B,nx]`, `cmd [B,H,nu]`, `traj [B,H+1,nx]`, ...).  The kernels use a structure-of-arrays layout with the
batch innermost (`[H][nu][B]`); the conversions live here.  PyTorch only provides device memory and the
current CUDA stream; every number is produced by the kernels of ilqc_b200/csrc.
"""
import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib

F64 = torch.float64


def parse_algo(algo: str):
    """Algorithm grammar of run_min_algo.check_nomenclature (algorithms/run_min_algo.py:267-292)."""
    if algo == "gd":
        return _lib.ALGO_GD, _lib.APPROX_LINQUAD, _lib.STEP_REGVAR
    if "newton" in algo:
        raise NotImplementedError(
            "newton_* (dense autodiff Hessian cross-check, algos_steps.py:46-102) is not part of the batched "
            "engine; use classic_quad_* for the structured Newton step")
    algo_type, approx, step_mode = algo.split("_")
    assert algo_type in ["classic", "ddp"]
    assert approx in ["linquad", "quad"]
    assert step_mode in ["dir", "reg", "dirvar", "regvar"]
    return ({"classic": _lib.ALGO_CLASSIC, "ddp": _lib.ALGO_DDP}[algo_type],
            {"linquad": _lib.APPROX_LINQUAD, "quad": _lib.APPROX_QUAD}[approx],
            {"dir": _lib.STEP_DIR, "reg": _lib.STEP_REG, "dirvar": _lib.STEP_DIRVAR, "regvar": _lib.STEP_REGVAR}[step_mode])


def default_stepsize(algo: str) -> float:
    """run_min_algo.py:59: 100 for the regularised step modes, 1 otherwise ('gd' parses as 'dir')."""
    if algo == "gd":
        return 1.0
    return 100.0 if "reg" in algo.split("_")[-1] else 1.0


@dataclass
class SolveResult:
    cmd: torch.Tensor          # [B,H,nu] final controls
    stepsize: torch.Tensor     # [B] optimiser state (optim_aux_vars['stepsize'])
    cost: torch.Tensor         # [B,n_iter+1]  metrics['cost'] (NaN-padded once a problem has stopped)
    norm_grad_obj: torch.Tensor  # [B,n_iter+1]
    stepsize_reported: torch.Tensor  # [B,n_iter+1] metrics['stepsize']
    n_done: torch.Tensor       # [B] iterations executed
    status: torch.Tensor       # [B] ILQC_RUNNING / CONVERGED / DIVERGED / STUCK / NO_DIRECTION
    ls_trips: Optional[torch.Tensor] = None  # [B,n_iter] line-search trips


class Engine:
    def __init__(self, lib, device):
        self.lib = lib
        self.device = torch.device(device)
        self._tables = {}
        self._workspace = None

    # ---- plumbing --------------------------------------------------------------------------------------
    def _stream(self):
        if self.device.type == "cuda":
            return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        return C.c_void_p(0)

    def _dev(self, x, dtype=F64):
        return torch.as_tensor(x, dtype=dtype).detach().to(self.device).contiguous()

    def empty(self, *shape, dtype=F64):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    @staticmethod
    def _ptr(t):
        return C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())

    def to_soa(self, x):
        """[B, ...] -> [prod(...), B] contiguous."""
        B = x.shape[0]
        return x.reshape(B, -1).t().contiguous()

    @staticmethod
    def from_soa(y, B, *shape):
        return y.reshape(-1, B).t().reshape(B, *shape).contiguous()

    def track_tables(self, env):
        tr = env.track
        if tr is None:
            return None, None
        key = tr.name
        if key not in self._tables:
            self._tables[key] = (self._dev(tr.knots), self._dev(tr.coef.reshape(-1)))
        return self._tables[key]

    def make_desc(self, env, horizon, B, t0=None, weights=None):
        """ModelDesc + the tensors that must stay alive while it is used."""
        d = env.model_desc(horizon)
        keep = []
        knots, coef = self.track_tables(env)
        if knots is not None:
            d.knots, d.coef = knots.data_ptr(), coef.data_ptr()
            keep += [knots, coef]
        if t0 is not None:
            if isinstance(t0, (int, np.integer)):
                d.t0 = int(t0)
            else:
                t0t = self._dev(t0, torch.int32).reshape(B)
                d.pp_t0 = t0t.data_ptr()
                keep.append(t0t)
        if weights is not None:
            w = self.to_soa(self._dev(weights).reshape(B, 3))
            d.pp_weights = w.data_ptr()
            keep.append(w)
        return d, keep

    def dims(self, d):
        nx, nu = C.c_int32(), C.c_int32()
        _lib.check(self.lib.ilqc_dims(C.byref(d), C.byref(nx), C.byref(nu)), "ilqc_dims")
        return nx.value, nu.value

    def _x0(self, env, x0, B=None):
        if x0 is None:
            x0 = env.init_state.reshape(1, -1)
            if B is not None:
                x0 = x0.expand(B, -1)
        return self._dev(x0)

    # ---- forward.py ------------------------------------------------------------------------------------
    def rollout_cost(self, env, x0, cmd, t0=None, weights=None):
        """DiffEnv.forward(cmd) for a batch: traj [B,H+1,nx], cost_t [B,H+1], total [B]."""
        cmd = self._dev(cmd)
        B, H, nu = cmd.shape
        x0 = self._x0(env, x0, B)
        d, keep = self.make_desc(env, H, B, t0, weights)
        nx, nu_ = self.dims(d)
        assert nu == nu_ and x0.shape == (B, nx)
        x0s, cmds = self.to_soa(x0), self.to_soa(cmd)
        traj, cost_t, total = self.empty((H + 1) * nx, B), self.empty(H + 1, B), self.empty(B)
        _lib.check(self.lib.ilqc_rollout_cost(C.byref(d), B, self._ptr(x0s), self._ptr(cmds), self._ptr(traj),
                                              self._ptr(cost_t), self._ptr(total), self._stream()), "ilqc_rollout_cost")
        return self.from_soa(traj, B, H + 1, nx), self.from_soa(cost_t, B, H + 1), total

    def linearize(self, env, x0, cmd, t0=None, weights=None):
        """Packed expansion (engine layout): lin [H*stride, B], traj [(H+1)*nx, B], total [B]."""
        cmd = self._dev(cmd)
        B, H, nu = cmd.shape
        x0 = self._x0(env, x0, B)
        d, keep = self.make_desc(env, H, B, t0, weights)
        nx, _ = self.dims(d)
        stride = self.lib.ilqc_lin_stride(C.byref(d))
        x0s, cmds = self.to_soa(x0), self.to_soa(cmd)
        lin, traj, total = self.empty(H * stride, B), self.empty((H + 1) * nx, B), self.empty(B)
        _lib.check(self.lib.ilqc_linearize(C.byref(d), B, self._ptr(x0s), self._ptr(cmds), self._ptr(lin),
                                           self._ptr(traj), self._ptr(total), self._stream()), "ilqc_linearize")
        return dict(lin=lin, traj=traj, total=total, x0=x0s, cmd=cmds, desc=d, keep=keep, B=B, H=H, nx=nx, nu=nu)

    def expand_dense(self, env, x0, cmd, t0=None, weights=None, lam=None):
        """Dense expansion in the reference's shapes; with lam [B,H,nx] also the dynamics Hessians."""
        cmd = self._dev(cmd)
        B, H, nu = cmd.shape
        x0 = self._x0(env, x0, B)
        d, keep = self.make_desc(env, H, B, t0, weights)
        nx, _ = self.dims(d)
        x0s, cmds = self.to_soa(x0), self.to_soa(cmd)
        A, Bm = self.empty(H * nx * nx, B), self.empty(H * nx * nu, B)
        p, P = self.empty((H + 1) * nx, B), self.empty((H + 1) * nx * nx, B)
        q, Q = self.empty(H * nu, B), self.empty(H * nu * nu, B)
        order = 1
        lams = Hxx = Hxu = Huu = None
        if lam is not None:
            order = 2
            lams = self.to_soa(self._dev(lam).reshape(B, H, nx))
            Hxx, Hxu, Huu = self.empty(H * nx * nx, B), self.empty(H * nx * nu, B), self.empty(H * nu * nu, B)
        _lib.check(self.lib.ilqc_expand_dense(C.byref(d), B, self._ptr(x0s), self._ptr(cmds), order, self._ptr(A),
                                              self._ptr(Bm), self._ptr(p), self._ptr(P), self._ptr(q), self._ptr(Q),
                                              self._ptr(lams), self._ptr(Hxx), self._ptr(Hxu), self._ptr(Huu),
                                              self._stream()), "ilqc_expand_dense")
        out = dict(A=self.from_soa(A, B, H, nx, nx), B=self.from_soa(Bm, B, H, nx, nu),
                   p=self.from_soa(p, B, H + 1, nx), P=self.from_soa(P, B, H + 1, nx, nx),
                   q=self.from_soa(q, B, H, nu), Q=self.from_soa(Q, B, H, nu, nu))
        if lam is not None:
            out.update(Hxx=self.from_soa(Hxx, B, H, nx, nx), Hxu=self.from_soa(Hxu, B, H, nx, nu),
                       Huu=self.from_soa(Huu, B, H, nu, nu))
        return out

    # ---- backward.py / rollout.py on a packed expansion -------------------------------------------------
    def backward(self, ex, mode, reg_ctrl, reg_state=0.0, L=1):
        """Backward pass for every (problem, candidate): reg_ctrl [B,L].  Returns gains (engine layout
        [H*gain_stride, B*L]), K [B,L,H,nu,nx], k [B,L,H,nu], jcst [B,L], feasible [B,L]."""
        B, H, nx, nu = ex["B"], ex["H"], ex["nx"], ex["nu"]
        d = ex["desc"]
        n_slots = B * L
        reg = self._dev(reg_ctrl).reshape(n_slots)
        prob = torch.arange(B, device=self.device, dtype=torch.int32).repeat_interleave(L).contiguous()
        gstride = self.lib.ilqc_gain_stride(C.byref(d))
        gains, jcst = self.empty(H * gstride, n_slots), self.empty(n_slots)
        feas = self.empty(n_slots, dtype=torch.int32)
        _lib.check(self.lib.ilqc_backward(C.byref(d), mode, B, n_slots, self._ptr(prob), self._ptr(reg),
                                          float(reg_state), self._ptr(ex["lin"]), self._ptr(ex["traj"]),
                                          self._ptr(ex["cmd"]), self._ptr(gains), self._ptr(jcst), self._ptr(feas),
                                          self._stream()), "ilqc_backward")
        g = gains.reshape(H, gstride, n_slots)
        K = g[:, :nu * nx].reshape(H, nu, nx, B, L).permute(3, 4, 0, 1, 2).contiguous()
        k = g[:, nu * nx:].reshape(H, nu, B, L).permute(2, 3, 0, 1).contiguous()
        return dict(gains=gains, K=K, k=k, jcst=jcst.reshape(B, L), feasible=feas.reshape(B, L).bool(), prob=prob, L=L)

    def rollout(self, ex, kind, step, bw=None, direction=None):
        """Candidate roll-outs: step [B,L]; bw = result of backward() (EXACT / LIN) or direction [B,H,nu] (GD).
        Returns diff [B,L,H,nu], newval [B,L], diffsq [B,L]."""
        B, H, nx, nu = ex["B"], ex["H"], ex["nx"], ex["nu"]
        d = ex["desc"]
        step = self._dev(step)
        L = step.numel() // B
        n_slots = B * L
        prob = torch.arange(B, device=self.device, dtype=torch.int32).repeat_interleave(L).contiguous()
        diff, newval, diffsq = self.empty(H * nu, n_slots), self.empty(n_slots), self.empty(n_slots)
        gains = bw["gains"] if bw is not None else None
        dirs = self.to_soa(self._dev(direction)) if direction is not None else None
        _lib.check(self.lib.ilqc_rollout(C.byref(d), kind, B, n_slots, self._ptr(prob), None,
                                         self._ptr(step.reshape(n_slots)), self._ptr(ex["x0"]), self._ptr(ex["cmd"]),
                                         self._ptr(ex["traj"]), self._ptr(ex["lin"]), self._ptr(gains), self._ptr(dirs),
                                         self._ptr(diff), self._ptr(newval), self._ptr(diffsq), self._stream()),
                   "ilqc_rollout")
        return (diff.reshape(H, nu, B, L).permute(2, 3, 0, 1).contiguous(), newval.reshape(B, L), diffsq.reshape(B, L))

    def grad_obj(self, ex):
        """grad_u sum(costs) [B,H,nu] and its norm [B]."""
        B, H, nu = ex["B"], ex["H"], ex["nu"]
        grad, gnorm = self.empty(H * nu, B), self.empty(B)
        _lib.check(self.lib.ilqc_grad_obj(C.byref(ex["desc"]), B, self._ptr(ex["lin"]), self._ptr(grad),
                                          self._ptr(gnorm), None, self._stream()), "ilqc_grad_obj")
        return self.from_soa(grad, B, H, nu), gnorm

    # ---- generic dense small systems --------------------------------------------------------------------
    def lq_backward_dense(self, A, Bm, p, P, q, Q, reg_ctrl, reg_state=0.0, R=None):
        """A [B,H,nx,nx], Bm [B,H,nx,nu], p [B,H+1,nx], P [B,H+1,nx,nx], q [B,H,nu], Q [B,H,nu,nu],
        R [B,H,nx,nu] or None, reg_ctrl [B,L] -> K [B,L,H,nu,nx], k [B,L,H,nu], jcst [B,L]."""
        A = self._dev(A)
        B, H, nx, _ = A.shape
        nu = Bm.shape[-1]
        reg = self._dev(reg_ctrl).reshape(B, -1)
        L = reg.shape[1]
        n = B * L
        s = [self.to_soa(self._dev(t)) for t in (A, Bm, p, P, q, Q)]
        Rs = self.to_soa(self._dev(R)) if R is not None else None
        K, k, jcst = self.empty(H * nu * nx, n), self.empty(H * nu, n), self.empty(n)
        _lib.check(self.lib.ilqc_lq_backward_dense(nx, nu, H, B, L, *[self._ptr(t) for t in s], self._ptr(Rs),
                                                   self._ptr(reg.reshape(n)), float(reg_state), self._ptr(K),
                                                   self._ptr(k), self._ptr(jcst), self._stream()),
                   "ilqc_lq_backward_dense")
        return (K.reshape(H, nu, nx, B, L).permute(3, 4, 0, 1, 2).contiguous(),
                k.reshape(H, nu, B, L).permute(2, 3, 0, 1).contiguous(), jcst.reshape(B, L))

    def rollout_lin_dense(self, A, Bm, K, k, step):
        """roll_out_lin on dense expansions: K [B,L,H,nu,nx], k [B,L,H,nu], step [B,L] -> diff [B,L,H,nu]."""
        A = self._dev(A)
        B, H, nx, _ = A.shape
        nu = Bm.shape[-1]
        K, k, step = self._dev(K), self._dev(k), self._dev(step).reshape(B, -1)
        L = step.shape[1]
        n = B * L
        As, Bs = self.to_soa(A), self.to_soa(self._dev(Bm))
        Ks = K.permute(2, 3, 4, 0, 1).reshape(H * nu * nx, n).contiguous()
        ks = k.permute(2, 3, 0, 1).reshape(H * nu, n).contiguous()
        diff = self.empty(H * nu, n)
        _lib.check(self.lib.ilqc_rollout_lin_dense(nx, nu, H, B, L, self._ptr(As), self._ptr(Bs), self._ptr(Ks),
                                                   self._ptr(ks), self._ptr(step.reshape(n).contiguous()),
                                                   self._ptr(diff), self._stream()), "ilqc_rollout_lin_dense")
        return diff.reshape(H, nu, B, L).permute(2, 3, 0, 1).contiguous()

    # ---- run_min_algo -------------------------------------------------------------------------------------
    def algo_desc(self, algo, n_candidates=1, line_search=True, compute_grad_norm=True, max_ls_trips=0, scheduler=0,
                  handle_bad_dir="check_total_value", expansion=0, cost0=None):
        a = _lib.AlgoDesc()
        a.algo_type, a.approx, a.step_mode = parse_algo(algo)
        a.n_candidates = int(n_candidates)
        a.line_search = int(bool(line_search))
        a.compute_grad_norm = int(bool(compute_grad_norm))
        a.max_ls_trips = int(max_ls_trips)
        a.scheduler = int(scheduler)
        if handle_bad_dir not in _lib.BAD_DIR:   # 'modify_hess' is dead code in the reference (removed torch.eig)
            raise NotImplementedError(handle_bad_dir)
        a.handle_bad_dir = _lib.BAD_DIR[handle_bad_dir]
        a.expansion = int(expansion)
        if cost0 is not None:      # restart: the original metrics['cost'][0] per problem (kept alive by the caller)
            a.pp_cost0 = cost0.data_ptr()
        return a

    def workspace(self, nbytes):
        if self._workspace is None or self._workspace.numel() < nbytes:
            self._workspace = None
            self._workspace = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return self._workspace

    def solve(self, env, algo="ddp_linquad_reg", max_iter=10, x0=None, cmd0=None, stepsize=None, batch=None,
              weights=None, t0=None, horizon=None, n_candidates=128, line_search=True, compute_grad_norm=True,
              record_trips=False, max_ls_trips=0, scheduler=0, handle_bad_dir="check_total_value", expansion=0,
              cost0=None):
        """run_min_algo for a batch of problems of one family (device tensors in / out).  cost0 [B]: the first cost
        of an earlier call this one continues (restart, run_min_algo(past_metrics=...))."""
        if cmd0 is not None:
            cmd0 = self._dev(cmd0)
            B, H, nu = cmd0.shape
        else:
            B = batch if batch is not None else (1 if x0 is None else torch.as_tensor(x0).reshape(-1, env.dim_state).shape[0])
            H = env.horizon if horizon is None else horizon
        x0 = self._x0(env, x0, B).reshape(B, -1)
        d, keep = self.make_desc(env, H, B, t0, weights)
        nx, nu = self.dims(d)
        cost0 = None if cost0 is None else self._dev(cost0).reshape(B)
        a = self.algo_desc(algo, n_candidates, line_search, compute_grad_norm, max_ls_trips, scheduler, handle_bad_dir, expansion,
                           cost0)
        cmds = self.to_soa(cmd0) if cmd0 is not None else torch.zeros(H * nu, B, dtype=F64, device=self.device)
        if stepsize is None:
            stepsize = default_stepsize(algo)
        step = (torch.full((B,), float(stepsize), dtype=F64, device=self.device)
                if not torch.is_tensor(stepsize) else self._dev(stepsize).reshape(B).clone())
        x0s = self.to_soa(x0)
        n = max_iter + 1
        cost, gnorm, srep = self.empty(n, B), self.empty(n, B), self.empty(n, B)
        n_done, status = self.empty(B, dtype=torch.int32), self.empty(B, dtype=torch.int32)
        trips = self.empty(max(max_iter, 1), B, dtype=torch.int32) if record_trips else None
        nbytes = self.lib.ilqc_solve_workspace_bytes(C.byref(d), C.byref(a), B)
        ws = self.workspace(nbytes)
        _lib.check(self.lib.ilqc_solve(C.byref(d), C.byref(a), B, self._ptr(x0s), self._ptr(cmds), self._ptr(step),
                                       max_iter, self._ptr(cost), self._ptr(gnorm), self._ptr(srep), self._ptr(n_done),
                                       self._ptr(status), self._ptr(trips), self._ptr(ws), C.c_size_t(nbytes),
                                       self._stream()), "ilqc_solve")
        return SolveResult(cmd=self.from_soa(cmds, B, H, nu), stepsize=step, cost=cost.t().contiguous(),
                           norm_grad_obj=gnorm.t().contiguous(), stepsize_reported=srep.t().contiguous(),
                           n_done=n_done, status=status,
                           ls_trips=trips[:max_iter].t().contiguous() if record_trips else None)

    def solve_host(self, env, algo, max_iter, x0_host, cmd_host, stepsize_host, weights=None, t0=None,
                   n_candidates=128, compute_grad_norm=True, scheduler=0, cmd0_host="same"):
        """Same solve through ilqc_solve_host: HOST (pinned) tensors in the reference layout, copies inside.
        cmd_host [B,H,nu] receives the controls; cmd0_host = initial controls ("same": cmd_host in place,
        None: zeros like run_min_algo(prev_cmd=None), nothing is copied to the device for them)."""
        if isinstance(cmd0_host, str):
            cmd0_host = cmd_host
        B, H, nu = cmd_host.shape
        d, keep = self.make_desc(env, H, B, t0, weights)
        a = self.algo_desc(algo, n_candidates, True, compute_grad_norm, 0, scheduler)
        n = max_iter + 1
        pin = self.device.type == "cuda"
        cost = torch.empty(B, n, dtype=F64, pin_memory=pin)
        gnorm = torch.empty(B, n, dtype=F64, pin_memory=pin)
        srep = torch.empty(B, n, dtype=F64, pin_memory=pin)
        n_done = torch.empty(B, dtype=torch.int32, pin_memory=pin)
        status = torch.empty(B, dtype=torch.int32, pin_memory=pin)
        nbytes = self.lib.ilqc_solve_host_workspace_bytes(C.byref(d), C.byref(a), B, max_iter)
        ws = self.workspace(nbytes)
        _lib.check(self.lib.ilqc_solve_host(C.byref(d), C.byref(a), B, self._ptr(x0_host), self._ptr(cmd0_host),
                                            self._ptr(cmd_host), self._ptr(stepsize_host), max_iter, self._ptr(cost), self._ptr(gnorm),
                                            self._ptr(srep), self._ptr(n_done), self._ptr(status), self._ptr(ws),
                                            C.c_size_t(nbytes), self._stream()), "ilqc_solve_host")
        return SolveResult(cmd=cmd_host, stepsize=stepsize_host, cost=cost, norm_grad_obj=gnorm,
                           stepsize_reported=srep, n_done=n_done, status=status)

    def mpc(self, env, x0, full_horizon, window_size, overlap, max_iter=10, algo="ddp_linquad_reg", weights=None,
            n_candidates=128, optim_on_full_window=False, keep_applying_last_ctrl=True, max_windows=0, expansion=0):
        """ilqc_mpc: B concurrent closed-loop episodes (run_mpc, mpc_pipeline.py:87-115), window loop inside the call.
        Returns (cmd [B,T,nu], cost [B,W,max_iter+1], status [B,W], n_done [B,W])."""
        x0 = self._dev(x0)
        B = x0.shape[0]
        d, keep = self.make_desc(env, env.horizon, B, env.init_time_iter, weights)
        nx, nu = self.dims(d)
        a = self.algo_desc(algo, n_candidates, True, True, 0, 0, "check_total_value", expansion)
        W = self.lib.ilqc_mpc_num_windows(int(full_horizon), int(window_size), int(overlap))
        if W <= 0:
            raise _lib.IlqcError("ilqc_mpc_num_windows: invalid window geometry")
        if max_windows:
            W = min(W, int(max_windows))
        nbytes = self.lib.ilqc_mpc_workspace_bytes(C.byref(d), C.byref(a), B, int(full_horizon), int(window_size), int(overlap),
                                                   int(max_iter), int(bool(optim_on_full_window)))
        if nbytes == 0:
            raise _lib.IlqcError("ilqc_mpc_workspace_bytes: invalid arguments")
        ws = self.workspace(nbytes)
        x0s = self.to_soa(x0)
        cmd = torch.zeros(int(full_horizon) * nu, B, dtype=F64, device=self.device)
        cost = self.empty(W, max_iter + 1, B)
        status, n_done = self.empty(W, B, dtype=torch.int32), self.empty(W, B, dtype=torch.int32)
        n_rows = C.c_int32(0)
        _lib.check(self.lib.ilqc_mpc(C.byref(d), C.byref(a), B, self._ptr(x0s), int(full_horizon), int(window_size), int(overlap),
                                     int(max_iter), int(bool(optim_on_full_window)), int(bool(keep_applying_last_ctrl)),
                                     int(max_windows or 0), self._ptr(cmd), C.byref(n_rows), self._ptr(cost), self._ptr(status),
                                     self._ptr(n_done), self._ptr(ws), C.c_size_t(nbytes), self._stream()), "ilqc_mpc")
        T = n_rows.value
        return (self.from_soa(cmd[:T * nu], B, T, nu), cost.permute(2, 0, 1).contiguous(), status.t().contiguous(),
                n_done.t().contiguous())

    # ---- dense diagnostics of the reference ------------------------------------------------------------------
    def newton_dense(self, env, x0, cmd, reg, t0=None, weights=None, want_hess=False):
        """newton_oracle's linear algebra (algos_steps.py:46-102) for a batch: dir [B,H,nu] = -(Hess + reg I)^-1 grad,
        opt [B] = dir.grad/2 (+ grad [B,H,nu] and, with want_hess, the dense Hessian [B,H nu,H nu])."""
        cmd = self._dev(cmd)
        B, H, nu = cmd.shape
        x0 = self._x0(env, x0, B)
        d, keep = self.make_desc(env, H, B, t0, weights)
        nbytes = self.lib.ilqc_newton_dense_workspace_bytes(C.byref(d), B)
        if nbytes == 0:
            raise _lib.IlqcError("ilqc_newton_dense_workspace_bytes: unsupported dimensions")
        ws = self.workspace(nbytes)
        x0s, cmds = self.to_soa(x0), self.to_soa(cmd)
        regs = self._dev(reg).reshape(B)
        n = H * nu
        dirs, opt, grad = self.empty(n, B), self.empty(B), self.empty(n, B)
        hess = self.empty(B, n, n) if want_hess else None
        _lib.check(self.lib.ilqc_newton_dense(C.byref(d), B, self._ptr(x0s), self._ptr(cmds), self._ptr(regs), self._ptr(dirs),
                                              self._ptr(opt), self._ptr(grad), self._ptr(hess), self._ptr(ws), C.c_size_t(nbytes),
                                              self._stream()), "ilqc_newton_dense")
        out = (self.from_soa(dirs, B, H, nu), opt, self.from_soa(grad, B, H, nu))
        return out + (hess,) if want_hess else out[:2]

    def traj_jacobian(self, env, x0, cmd, t0=None, weights=None):
        """d traj[1:] / d cmd for a batch: [B, H nx, H nu] (compute_min_sing_eigval_jac, run_min_algo.py:103-124)."""
        cmd = self._dev(cmd)
        B, H, nu = cmd.shape
        x0 = self._x0(env, x0, B)
        d, keep = self.make_desc(env, H, B, t0, weights)
        nx, _ = self.dims(d)
        nbytes = self.lib.ilqc_traj_jacobian_workspace_bytes(C.byref(d), B)
        if nbytes == 0:
            raise _lib.IlqcError("ilqc_traj_jacobian_workspace_bytes: unsupported dimensions")
        ws = self.workspace(nbytes)
        x0s, cmds = self.to_soa(x0), self.to_soa(cmd)
        jac = self.empty(B, H * nx, H * nu)
        _lib.check(self.lib.ilqc_traj_jacobian(C.byref(d), B, self._ptr(x0s), self._ptr(cmds), self._ptr(jac), self._ptr(ws),
                                               C.c_size_t(nbytes), self._stream()), "ilqc_traj_jacobian")
        return jac

    def fastmath_eval(self, fn, a, b=None):
        """Self-test of csrc/fastmath.cuh: fn in {sincos, atan2, atan, div, rcp, sqrt_rsqrt}, elementwise."""
        code = ["sincos", "atan2", "atan", "div", "rcp", "sqrt_rsqrt"].index(fn)
        a = a.to(self.device, F64).contiguous()
        b = b.to(self.device, F64).contiguous() if b is not None else None
        o0, o1 = torch.empty_like(a), torch.empty_like(a)
        _lib.check(self.lib.ilqc_fastmath_eval(code, a.numel(), self._ptr(a), self._ptr(b) if b is not None else None,
                                               self._ptr(o0), self._ptr(o1), self._stream()), "ilqc_fastmath_eval")
        return o0, o1

    def fp64_peak(self, iters=4096):
        out = C.c_double()
        _lib.check(self.lib.ilqc_fp64_peak(int(iters), C.byref(out), self._stream()), "ilqc_fp64_peak")
        return out.value


_default = None


def default_engine():
    """The product engine: CUDA library on the current CUDA device.  No CPU fallback."""
    global _default
    if _default is None:
        if not torch.cuda.is_available():
            raise RuntimeError("ilqc_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        _default = Engine(_lib.cuda_lib(), torch.device("cuda", torch.cuda.current_device()))
    return _default
