"""Time the hot kernels of the bicycle workload through the C ABI for several thread counts and library
variants (run on the GPU box).  python tools/tune_kernels.py > gpurun_out/tuning.txt"""
import ctypes as C
import glob
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ilqc_b200 import _lib  # noqa: E402
from ilqc_b200.engine import Engine  # noqa: E402
from bench import make_batch  # noqa: E402


def time_call(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    B = 32768
    env, x0, w = make_batch(B, 1236)
    libs = {"default": _lib.CUDA_LIB_PATH}
    if not os.environ.get("ILQC_TUNE_DEFAULT_ONLY"):
        for p in sorted(glob.glob(os.path.join(ROOT, "ilqc_b200", "build", "variants", "libilqc_*.so"))):
            libs[os.path.basename(p)[8:-3]] = p
    print("variant, kernel, n_threads, ms, threads_per_ms")
    for tag, path in libs.items():
        eng = Engine(_lib.bind(path), torch.device("cuda", 0))
        # a realistic iterate: 2 solver iterations from zero
        r = eng.solve(env, "ddp_linquad_reg", 2, x0=x0, weights=w)
        ex = eng.linearize(env, x0, r.cmd, weights=w)
        d = ex["desc"]
        st = eng._stream()
        gstride = eng.lib.ilqc_gain_stride(C.byref(d))
        H, nx, nu = ex["H"], ex["nx"], ex["nu"]
        t = time_call(lambda: eng.lib.ilqc_linearize(C.byref(d), B, eng._ptr(ex["x0"]), eng._ptr(ex["cmd"]), eng._ptr(ex["lin"]),
                                                     eng._ptr(ex["traj"]), eng._ptr(ex["total"]), st))
        print(f"{tag}, linearize, {B}, {t:.3f}, {B / t:.0f}", flush=True)
        # (active problems, candidates per problem): full waves and the straggler rounds of the line search
        for nprob, L, order in [(B, 1, "p"), (B, 2, "p"), (B, 2, "c"), (B, 4, "p"), (B, 4, "c"), (B // 2, 4, "p"), (B // 2, 4, "c"),
                                (B // 4, 9, "p"), (B // 4, 9, "c"), (B // 16, 32, "p"), (B // 16, 32, "c"),
                                (B // 128, 32, "p"), (B // 128, 32, "c"), (32, 32, "p"), (32, 32, "c")]:
            n = nprob * L
            sel = torch.randperm(B, device="cuda")[:nprob].sort().values.to(torch.int32)
            regs = 1.0 / (100.0 * 0.5 ** torch.arange(L, device="cuda", dtype=torch.float64))
            if order == "p":   # problem-major: the candidates of a problem are adjacent threads
                prob = sel.repeat_interleave(L).contiguous()
                reg = regs.repeat(nprob).contiguous()
            else:              # candidate-major: adjacent threads are adjacent problems
                prob = sel.repeat(L).contiguous()
                reg = regs.repeat_interleave(nprob).contiguous()
            gains, jcst = eng.empty(H * gstride, n), eng.empty(n)
            feas = eng.empty(n, dtype=torch.int32)
            step = torch.ones(n, device="cuda", dtype=torch.float64)
            diff, newval, dsq = eng.empty(H * nu, n), eng.empty(n), eng.empty(n)
            tb = time_call(lambda: eng.lib.ilqc_backward(C.byref(d), 0, B, n, eng._ptr(prob), eng._ptr(reg), 0.0, eng._ptr(ex["lin"]),
                                                         eng._ptr(ex["traj"]), eng._ptr(ex["cmd"]), eng._ptr(gains), eng._ptr(jcst),
                                                         eng._ptr(feas), st))
            tr = time_call(lambda: eng.lib.ilqc_rollout(C.byref(d), 0, B, n, eng._ptr(prob), None, eng._ptr(step), eng._ptr(ex["x0"]),
                                                        eng._ptr(ex["cmd"]), eng._ptr(ex["traj"]), eng._ptr(ex["lin"]), eng._ptr(gains),
                                                        None, eng._ptr(diff), eng._ptr(newval), eng._ptr(dsq), st))
            print(f"{tag}, backward, {nprob}x{L}{order}={n}, {tb:.3f}, {n / tb:.0f}", flush=True)
            print(f"{tag}, rollout, {nprob}x{L}{order}={n}, {tr:.3f}, {n / tr:.0f}", flush=True)
            del gains, diff
        del eng
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
