"""GPU box: warp-cooperative backward pass (csrc/coop.cuh) against the thread-per-candidate kernel, by launch size, then
the whole headline solve and an MPC window workload with different switch-over points (ILQC tuning knob coop_max)."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_batch  # noqa: E402
from ilqc_b200.engine import default_engine  # noqa: E402
eng = default_engine()


def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


env, x0, w = make_batch(32768, 1236)
x0, w = x0.cuda(), w.cuda()
print("# backward pass alone: candidates, thread-per-candidate ms, cooperative ms (bicycle H=100, linear-quadratic mode)")
for nprob, L in ((1, 1), (16, 4), (64, 16), (256, 8), (1024, 4), (2048, 4), (4096, 4), (8192, 4), (16384, 4), (32768, 4)):
    g = torch.Generator().manual_seed(5)
    cmd = 0.3 * torch.randn(nprob, env.horizon, env.dim_ctrl, generator=g, dtype=torch.float64).cuda()
    ex = eng.linearize(env, x0[:nprob], cmd, weights=w[:nprob])
    reg = (10.0 ** torch.linspace(-2, 4, L, dtype=torch.float64)).reshape(1, L).repeat(nprob, 1).cuda()
    out = []
    for cm in (0, 1 << 40):
        eng.lib.ilqc_set_tuning(b"coop_max", cm)
        out.append(timed(lambda: eng.lib.ilqc_backward is not None and eng.backward(ex, 0, reg, L=L)))
    print(nprob * L, "%.3f %.3f" % tuple(out), flush=True)
print("# headline solve (32768 problems x 20 iterations), ms per solve by coop_max")
for cm in (0, 2048, 8192, 16384, 24576, 49152):
    eng.lib.ilqc_set_tuning(b"coop_max", cm)
    print(cm, "%.2f" % timed(lambda: eng.solve(env, "ddp_linquad_reg", 20, x0=x0, weights=w), n=3), flush=True)
print("# warm-started small batch (4096 problems, 10 iterations from a converged start), ms per solve by coop_max")
r = eng.solve(env, "ddp_linquad_reg", 12, x0=x0[:4096], weights=w[:4096])
for cm in (0, 2048, 8192, 16384, 24576, 49152, 1 << 40):
    eng.lib.ilqc_set_tuning(b"coop_max", cm)
    print(cm, "%.2f" % timed(lambda: eng.solve(env, "ddp_linquad_reg", 10, x0=x0[:4096], cmd0=r.cmd, weights=w[:4096]), n=3), flush=True)
eng.lib.ilqc_set_tuning(b"coop_max", -1)
