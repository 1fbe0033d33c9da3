"""GPU box: per-kernel-class time of the second-order algorithm on the cfg4 slice."""
import ctypes as C
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_batch, CLASSES  # noqa: E402
from ilqc_b200 import _lib  # noqa: E402
from ilqc_b200.engine import Engine, default_engine  # noqa: E402
eng = Engine(_lib.bind(os.environ['ILQC_LIB']), torch.device('cuda', 0)) if os.environ.get('ILQC_LIB') else default_engine()
env, x0, w = make_batch(32768, 1236)
x0, w = x0.cuda(), w.cuda()
for algo, it in (("ddp_quad_reg", 10),):
    eng.solve(env, algo, 2, x0=x0, weights=w)
    torch.cuda.synchronize()
    ms, n, u = (C.c_double * 8)(), (C.c_int64 * 8)(), (C.c_int64 * 8)()
    eng.lib.ilqc_profile_enable(1)
    eng.lib.ilqc_profile_read(ms, n, u, 8)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = eng.solve(env, algo, it, x0=x0, weights=w, record_trips=True)
    e1.record()
    torch.cuda.synchronize()
    eng.lib.ilqc_profile_read(ms, n, u, 8)
    eng.lib.ilqc_profile_enable(0)
    print(algo, "total ms %.1f" % e0.elapsed_time(e1), {c: (round(ms[i], 1), n[i], u[i]) for i, c in enumerate(CLASSES)},
          "mean trips", r.ls_trips.float().mean(0).tolist()[:10], flush=True)
