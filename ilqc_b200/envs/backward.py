"""lin_quad_backward / quad_backward(_ddp/_newton) with the reference's signatures (envs/backward.py:6-150),
evaluated by the CUDA kernels.  `traj`, `costs` are what DiffEnv.forward(cmd, approx=...) returned; the
expansion is redone on the device from (env, x0, cmd) recorded on traj[0] (the reference keeps it as autograd
graphs on the same objects)."""
import torch

from .. import _lib


def _run(traj, costs, reg_ctrl, reg_state, handle_bad_dir, mode):
    if handle_bad_dir not in _lib.BAD_DIR:   # 'modify_hess' needs the removed torch.eig in the reference
        raise NotImplementedError("handle_bad_dir=%r" % handle_bad_dir)
    info = getattr(traj[0], "_ilqc", None)
    if info is None:
        raise ValueError("traj must come from ilqc_b200 DiffEnv.forward(cmd, approx='linquad'|'quad')")
    env = info["env"]
    eng = env._engine()
    H = info["cmd"].shape[0]
    ex = eng.linearize(env, info["x0"], info["cmd"].reshape(1, H, -1), t0=info["t0"])
    bw = eng.backward(ex, mode | (_lib.BAD_DIR[handle_bad_dir] << 4),
                      torch.tensor([[float(reg_ctrl)]], dtype=torch.float64), reg_state=float(reg_state))
    K, k = bw["K"][0, 0].cpu(), bw["k"][0, 0].cpu()
    gains = [dict(gain_ctrl=K[t], offset_ctrl=k[t]) for t in range(H)]
    jcst = bw["jcst"][0, 0].cpu()
    feasible = bool(bw["feasible"][0, 0])
    if not feasible and handle_bad_dir == "flag":
        # the reference stops at the first non-convex step (backward.py:55-56): gains of the later steps only,
        # next_jcst = None.  Callers only test `feasible` (algos_steps.py:140-147, :166-168)
        jcst = None
    return gains, jcst, feasible


def lin_quad_backward(traj, costs, reg_ctrl=0.0, reg_state=0.0, handle_bad_dir="check_total_value"):
    return _run(traj, costs, reg_ctrl, reg_state, handle_bad_dir, _lib.BWD_LINQUAD)


def quad_backward(traj, costs, reg_ctrl=0.0, reg_state=0.0, handle_bad_dir="check_total_value", mode="newton"):
    return _run(traj, costs, reg_ctrl, reg_state, handle_bad_dir,
                {"newton": _lib.BWD_QUAD_NEWTON, "ddp": _lib.BWD_QUAD_DDP}[mode])


def quad_backward_ddp(traj, costs, reg_ctrl=0.0, reg_state=0.0, handle_bad_dir="check_total_value"):
    return quad_backward(traj, costs, reg_ctrl, reg_state, handle_bad_dir, mode="ddp")


def quad_backward_newton(traj, costs, reg_ctrl=0.0, reg_state=0.0, handle_bad_dir="check_total_value"):
    return quad_backward(traj, costs, reg_ctrl, reg_state, handle_bad_dir, mode="newton")
