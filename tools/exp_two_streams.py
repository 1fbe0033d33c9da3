"""GPU box experiment: does solving two half-batches concurrently (two host threads, two CUDA streams, each the
plain lock-step driver) beat one full batch?  The expansion / backward kernels are latency-bound at 8 warps per SM,
so a second stream can fill the idle issue slots."""
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_batch, ALGO  # noqa: E402
from ilqc_b200 import _lib  # noqa: E402
from ilqc_b200.engine import Engine  # noqa: E402

B, ITERS = 32768, 20
env, x0, w = make_batch(B, 1236)
dev = torch.device("cuda", 0)
x0d, wd = x0.to(dev), w.to(dev)


def run(parts, reps=3):
    engines = [Engine(_lib.cuda_lib(), dev) for _ in range(parts)]
    streams = [torch.cuda.Stream(dev) for _ in range(parts)]
    chunk = B // parts

    def work(i):
        with torch.cuda.stream(streams[i]):
            sl = slice(i * chunk, (i + 1) * chunk)
            r = engines[i].solve(env, ALGO, ITERS, x0=x0d[sl], weights=wd[sl], n_candidates=128)
            streams[i].synchronize()
            return r

    def once():
        th = [threading.Thread(target=work, args=(i,)) for i in range(parts)]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    once()
    ts = [once() for _ in range(reps)]
    best = min(ts)
    print(f"parts={parts} best {best * 1e3:.1f} ms  -> {B * ITERS / best / 1e6:.2f} M iterations/s  all: {[round(t * 1e3, 1) for t in ts]}", flush=True)


for parts in (1, 2, 3, 4):
    run(parts)
