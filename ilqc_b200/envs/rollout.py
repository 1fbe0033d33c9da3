"""roll_out_lin / roll_out_exact with the reference's signatures (envs/rollout.py:8-88) on the CUDA kernels."""
import torch

from .. import _lib


def _gains_to_slots(eng, gains, nx, nu):
    H = len(gains)
    K = torch.stack([torch.as_tensor(g["gain_ctrl"], dtype=torch.float64).reshape(nu, nx) for g in gains])
    k = torch.stack([torch.as_tensor(g["offset_ctrl"], dtype=torch.float64).reshape(nu) for g in gains])
    rec = torch.cat((K.reshape(H, nu * nx), k), dim=1)          # [H, gain_stride], one slot
    return dict(gains=eng._dev(rec.reshape(H * (nu * nx + nu), 1)))


def _roll(traj, gains, stepsize, kind, cmd=None):
    info = getattr(traj[0], "_ilqc", None)
    if info is None:
        raise ValueError("traj must come from ilqc_b200 DiffEnv.forward(cmd, approx='linquad'|'quad')")
    env = info["env"]
    eng = env._engine()
    H = info["cmd"].shape[0]
    base_cmd = info["cmd"] if cmd is None else torch.as_tensor(cmd, dtype=torch.float64).detach()
    ex = eng.linearize(env, info["x0"], info["cmd"].reshape(1, H, -1), t0=info["t0"])
    if cmd is not None:
        ex["cmd"] = eng.to_soa(eng._dev(base_cmd.reshape(1, H, -1)))
    bw = _gains_to_slots(eng, gains, env.dim_state, env.dim_ctrl)
    diff, _, _ = eng.rollout(ex, kind, torch.tensor([[float(stepsize)]], dtype=torch.float64), bw)
    return diff[0, 0].cpu()


def roll_out_lin(traj, gains, stepsize=1.0):
    return _roll(traj, gains, stepsize, _lib.ROLL_LIN)


def roll_out_exact(env, traj, gains, cmd, stepsize=1.0):
    return _roll(traj, gains, stepsize, _lib.ROLL_EXACT, cmd)
