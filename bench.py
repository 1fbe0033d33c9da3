#!/usr/bin/env python
"""Benchmark of the batched iLQR/DDP hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
    python bench.py --impl reference --gpus N --steps K ...   # CPU reference arm (oracle port, host cores)

Workload (config.workload): BASELINE.json configs[3] per-GPU slice = bicycle (Pacejka) car on the simple
track, contouring cost, RK4 constant-control, H=100, dt=0.01, `ddp_linquad_reg` (the algorithm the metric /
target is quoted on), 20 solver iterations from cmd=0, 32768 problems per GPU with seeded per-problem
vinit and cost weights.  A "step" is one full solve of the batch (ilqc_solve).  Weak scaling:
every rank solves its own 32768 problems; the only collective is the final all-gather of results.

One JSON line on stdout (rank 0).  metric = solver iterations per second = sum over problems of iterations
executed / time.  `value`: inputs resident in HBM.  `e2e`: through ilqc_solve_host with pinned HOST buffers,
H2D/D2H copies inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# stdout carries exactly one JSON line: NCCL's version banner / debug log go to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":   # (that level printf()s its banner to stdout)
    os.environ["NCCL_DEBUG"] = "WARN"

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ENV_CFG = dict(env="real_car", track="simple", cost="contouring", discretization="rk4_cst_ctrl", horizon=100, dt=0.01)
ALGO = "ddp_linquad_reg"
METRIC = "DDP solver iterations/s (problems x iterations, bicycle car H=100 ddp_linquad_reg)"
UNIT = "iterations/s"

# Frozen flop convention of SURVEY.md 8(d), cfg4 row (nx=8, nu=3, s=4): algorithmic flops per time step per slot
F_PHI, F_COST, F_LIN, F_BELL, F_ROLL = 356, 150, 6164, 4394, 418
NX, NU = 8, 3
FLOPS_PER_STEP = {  # kernel class -> algorithmic flops per (slot, time step)
    "rollout_cost": F_PHI + F_COST,          # forward_kernel<.,0>
    "linearize": F_LIN,                      # forward_kernel<.,1>
    "backward": F_BELL,                      # backward_kernel (dense-equivalent count of bell_step)
    "rollout_candidate": F_ROLL + F_COST,    # rollout_kernel
    "adjoint": 2 * (NX * NX + NX * NU),      # adjoint_kernel
}
CLASSES = ["rollout_cost", "linearize", "backward", "rollout_candidate", "adjoint"]
# EXECUTED FP64 flops per (slot, time step) = dadd + dmul + 2 dfma thread instructions counted by ncu
# (smsp__sass_thread_inst_executed_op_{dadd,dmul,dfma}_pred_on, profiles/r01_ncu_hot_kernels_v3.txt).  The convention
# above counts a transcendental or a division as ONE flop (each is 10-40 FP64 instructions) and the Riccati step
# as dense 8x8 algebra (the kernel skips the structural zeros), so both views are reported.
EXEC_FLOPS_PER_STEP = {"rollout_cost": 2000, "linearize": 4604, "backward": 1495, "rollout_candidate": 2115, "adjoint": 176}
# DRAM bytes per slot per launch of the same ncu capture (dram__bytes_read.sum + dram__bytes_write.sum of a
# 131072-slot launch = 4 candidates for each of 32768 problems; per-problem rows are shared by the candidates)
NCU_DRAM_BYTES_PER_SLOT = {"rollout_candidate": 26350, "backward": 31900, "linearize": 50100}
LIN_STRIDE, GAIN_STRIDE = 52, 27
BYTES_PER_SLOT = {  # algorithmic HBM bytes per slot per launch (H=100): reads + writes of the SoA arrays
    "rollout_cost": 8 * (NX + 100 * NU + 1),
    "linearize": 8 * (NX + 100 * NU + 100 * LIN_STRIDE + 101 * NX + 2),
    "backward": 8 * (100 * LIN_STRIDE + 100 * GAIN_STRIDE + 3),
    "rollout_candidate": 8 * (NX + 100 * NU + 101 * NX + 100 * GAIN_STRIDE + 100 * NU + 3),
    "adjoint": 8 * (100 * LIN_STRIDE + 100 * NU + 1),
}


def make_batch(B, seed, H=100):
    """x0 [B,8] and weights [B,3]: vinit_b = 1 + 0.01 xi, reg_* = (0.1, 10, 0.1) + 1e-3 xi (car.py:49-56)."""
    import torch
    from ilqc_b200.envs import make_env
    cfg = dict(ENV_CFG)
    cfg["horizon"] = H
    env = make_env(cfg)
    g = torch.Generator().manual_seed(seed)
    xi = torch.randn(B, 4, generator=g, dtype=torch.float64)
    x0 = env.init_state.reshape(1, -1).repeat(B, 1)
    x0[:, 3] = 1.0 + 0.01 * xi[:, 0]
    x0[:, 7] = x0[:, 3]
    w = torch.tensor([0.1, 10.0, 0.1], dtype=torch.float64).reshape(1, 3) + 1e-3 * xi[:, 1:]
    return env, x0, w


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe): ONE
    `nvidia-smi -lms 200` process, started before the warm-up (its start-up takes driver locks that stall
    kernel launches for 100-200 ms) and killed after; only the samples that arrive between mark() and
    __exit__ are summarised."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.all_samples, self.t_mark = index, [], [], None
        self.proc = None
        self.thread = threading.Thread(target=self.run, daemon=True)

    def run(self):
        try:
            for line in self.proc.stdout:
                row = [s.strip() for s in line.strip().split(",")]
                self.all_samples.append(row)
                if self.t_mark is not None:
                    self.samples.append(row)
        except Exception:
            pass

    def mark(self):
        """start of the timed region"""
        self.t_mark = time.time()

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.25)  # let the last sample of the region arrive
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            self.thread.join(timeout=5)
        if not self.samples:  # region shorter than one sampling period: the latest sample before it ended
            self.samples = self.all_samples[-1:]

    def summary(self):
        sm = sorted(float(s[0]) for s in self.samples if len(s) >= 7 and s[0].replace(".", "").isdigit())
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        reasons = set()
        for s in self.samples:
            if len(s) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": sorted(reasons),
                "n_samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# CPU reference arm: the oracle port of the reference's algorithm on the host cores
# ------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    seed, iters = args
    import torch
    torch.set_num_threads(1)
    from oracle import ilqc_oracle as orc
    env, x0, w = make_batch(1, seed)
    cfg = dict(ENV_CFG, reg_cont=float(w[0, 0]), reg_lag=float(w[0, 1]), reg_speed=float(w[0, 2]))
    oenv = orc.make_env(cfg)
    oenv.init_state = x0[0].clone()
    t0 = time.time()
    _, _, m = orc.run_min_algo(oenv, max_iter=iters, algo=ALGO)
    return len(m["cost"]) - 1, time.time() - t0


def cpu_reference_step(pool, n_proc, step_idx, iters=1):
    """One bounded sample: n_proc problems of the workload, `iters` solver iteration(s) each, one process per
    host core (torch.set_num_threads(1)), exactly as BASELINE.md section 4 plans it."""
    t0 = time.time()
    res = pool.map(_cpu_worker, [(1236 + 1000 * step_idx + i, iters) for i in range(n_proc)])
    wall = time.time() - t0
    return sum(r[0] for r in res), wall


def run_reference(args):
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_proc = min(os.cpu_count() or 1, args.cpu_procs)
    # bounded: every step costs ~30-60 s of wall (one reference-style iteration of a H=100 bicycle problem per
    # core); cap the number of steps so that the run ends within a few minutes and report what was timed
    steps = max(1, min(args.steps, args.ref_max_steps))
    warm = min(args.warmup, 1)
    ctx = mp.get_context("spawn")
    with ctx.Pool(n_proc) as pool:
        for i in range(warm):
            cpu_reference_step(pool, n_proc, 100 + i)
        its, wall = 0, 0.0
        for i in range(steps):
            a, b = cpu_reference_step(pool, n_proc, i)
            its += a
            wall += b
    value = its / wall
    sample = f"{n_proc} problems x 1 iteration per step (one process per core, 1 thread each), {steps} steps"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1e3 * wall / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "bicycle car RK4 H=100 dt=0.01 contouring, ddp_linquad_reg (BASELINE configs[3])",
                   "note": "CPU oracle port of the reference algorithm (torch autograd, float64); reference is pure "
                           "Python and cannot travel to the GPU box"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": n_proc, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
def run_ours(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    from ilqc_b200.engine import default_engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = default_engine()
    dev = eng.device
    B, iters, L = args.batch, args.iters, args.candidates
    env, x0, w = make_batch(B, 1236 + rank)
    x0d, wd = x0.to(dev), w.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    gathered = {}

    def step():
        r = eng.solve(env, ALGO, iters, x0=x0d, weights=wd, n_candidates=L, scheduler=args.scheduler)
        if world > 1:  # the path's only collective: all-gather of the results
            for name, t in (("cost", r.cost), ("cmd", r.cmd), ("stepsize", r.stepsize), ("status", r.status)):
                out = gathered.get(name)
                if out is None:
                    out = gathered[name] = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
                dist.all_gather_into_tensor(out, t.contiguous())
        return r

    lib = eng.lib
    ms_buf, n_buf, u_buf = (C.c_double * 8)(), (C.c_int64 * 8)(), (C.c_int64 * 8)()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        lib.ilqc_profile_enable(1)   # (enabled before the warm-up so that its event pool exists when timing starts)
        for _ in range(args.warmup):
            r = step()
            int(r.n_done.sum().item())   # (same host-side reduction as the timed loop: its kernel loads lazily)
        barrier()
        lib.ilqc_profile_read(ms_buf, n_buf, u_buf, 8)  # reset
        lib.ilqc_launch_count(1)
        clocks.mark()
        barrier()
        e0.record()
        its = 0
        step_ev = [e0]
        for _ in range(args.steps):
            r = step()
            its += int(r.n_done.sum().item())
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            step_ev.append(ev)
        e1.record()
        barrier()
    n_launches = int(lib.ilqc_launch_count(1))
    lib.ilqc_profile_read(ms_buf, n_buf, u_buf, 8)
    lib.ilqc_profile_enable(0)
    timed = {c: dict(ms=ms_buf[i], launches=int(n_buf[i]), slots=int(u_buf[i])) for i, c in enumerate(CLASSES)}
    per_step_ms = [round(step_ev[i].elapsed_time(step_ev[i + 1]), 2) for i in range(args.steps)]
    print(f"[bench] rank {rank} per-step ms: {per_step_ms}", file=sys.stderr, flush=True)
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    tot_its = torch.tensor([float(its)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot_its, op=dist.ReduceOp.SUM)
    elapsed_s = ms.item() * 1e-3
    value = tot_its.item() / elapsed_s

    # ---- end-to-end through the C ABI with HOST buffers (copies inside the timed region) ----
    # every step copies its inputs (initial states, per-problem cost weights, step sizes) from pinned host memory and
    # reads its results (controls, metrics, status) back; the initial controls are zeros = run_min_algo(prev_cmd=None),
    # created on the device like the reference creates them itself (run_min_algo.py:57)
    x0h = x0.clone().pin_memory()
    wh = w.clone().pin_memory()
    wd_e2e = torch.empty_like(wd)
    cmdh = torch.empty(B, 100, NU, dtype=torch.float64).pin_memory()
    steph = torch.empty(B, dtype=torch.float64).pin_memory()
    e2e_steps = max(1, min(args.steps, 5))

    def e2e_step():
        steph.fill_(100.0)
        wd_e2e.copy_(wh, non_blocking=True)
        return eng.solve_host(env, ALGO, iters, x0h, cmdh, steph, weights=wd_e2e, n_candidates=L, scheduler=args.scheduler,
                              cmd0_host=None)

    for _ in range(2):   # (two result sets alternate in the pinned-memory cache: warm both)
        rh = e2e_step()
        int(rh.n_done.sum().item())
    barrier()
    t0 = time.perf_counter()
    e2e_its = 0
    e2e_marks = [t0]
    for _ in range(e2e_steps):
        rh = e2e_step()
        e2e_its += int(rh.n_done.sum().item())
        e2e_marks.append(time.perf_counter())
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    print(f"[bench] rank {rank} e2e per-step ms: {[round(1e3 * (b - a), 2) for a, b in zip(e2e_marks, e2e_marks[1:])]}",
          file=sys.stderr, flush=True)
    e2e_tot = torch.tensor([float(e2e_its)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_tot, op=dist.ReduceOp.SUM)
    h2d = 8 * (B * NX + 3 * B + B)
    d2h = 8 * (B * 100 * NU + B + 3 * B * (iters + 1)) + 8 * B

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----
    # Per-kernel durations: CUDA events around every launch (on the launching stream) of one extra lock-step
    # solve of the same batch right after the timed region (`kernels`); the event times collected inside the
    # timed region itself are reported as `kernels_timed_region`.
    lib.ilqc_profile_enable(1)
    lib.ilqc_profile_read(ms_buf, n_buf, u_buf, 8)
    eng.solve(env, ALGO, iters, x0=x0d, weights=wd, n_candidates=L, scheduler=1)
    torch.cuda.synchronize()
    lib.ilqc_profile_read(ms_buf, n_buf, u_buf, 8)
    lib.ilqc_profile_enable(0)
    prof = {c: dict(ms=ms_buf[i], launches=int(n_buf[i]), slots=int(u_buf[i])) for i, c in enumerate(CLASSES)}
    dom = max(CLASSES, key=lambda c: prof[c]["ms"])
    kernel_ms_total = sum(p["ms"] for p in prof.values())
    peak_fp64 = eng.fp64_peak(2048)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    pd = prof[dom]
    flops = FLOPS_PER_STEP[dom] * 100 * pd["slots"]
    achieved = flops / (pd["ms"] * 1e-3) / 1e12 if pd["ms"] > 0 else 0.0
    bytes_ = BYTES_PER_SLOT[dom] * pd["slots"]
    exec_flops = EXEC_FLOPS_PER_STEP[dom] * 100 * pd["slots"]
    exec_tflops = exec_flops / (pd["ms"] * 1e-3) / 1e12 if pd["ms"] > 0 else 0.0
    traffic = NCU_DRAM_BYTES_PER_SLOT.get(dom)
    roofline = {"bound": "fp64", "kernel": dom, "achieved": achieved, "peak": peak_fp64 / 1e12, "unit": "TFLOP/s",
                "frac": achieved / (peak_fp64 / 1e12),
                "traffic": None if traffic is None else traffic * pd["slots"] / max(pd["launches"], 1),
                "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per slot (profiles/r01_ncu_hot_kernels_v3.txt) "
                                  "x average slots per launch",
                "algorithmic_bytes_per_launch_avg": bytes_ / max(pd["launches"], 1),
                "executed": {"achieved": exec_tflops, "unit": "TFLOP/s", "frac": exec_tflops / (peak_fp64 / 1e12),
                             "flops_per_slot_step": EXEC_FLOPS_PER_STEP[dom],
                             "note": "FP64 instructions actually executed (ncu dadd + dmul + 2 dfma per slot-step) / "
                                     "live CUDA-event time; the convention counts a transcendental as one flop"},
                "peak_source": "measured DFMA micro-benchmark (ilqc_fp64_peak) on this GPU; nominal 37 TFLOP/s",
                "flops_per_launch_avg": flops / max(pd["launches"], 1), "ms_per_launch_avg": pd["ms"] / max(pd["launches"], 1),
                "share_of_kernel_time": pd["ms"] / kernel_ms_total if kernel_ms_total else None,
                "hbm": {"achieved": bytes_ / (pd["ms"] * 1e-3) / 1e9 if pd["ms"] > 0 else 0.0, "peak": hbm_peak, "unit": "GB/s",
                        "frac": (bytes_ / (pd["ms"] * 1e-3) / 1e9) / hbm_peak if pd["ms"] > 0 else 0.0,
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s"},
                "convention": "SURVEY.md 8(d) frozen table, cfg4 row",
                "timing": "CUDA events around every launch of one lock-step solve of the same batch (kernels not overlapped)"}
    # whole-iteration algorithmic flops: F_iter formula with the measured candidates per iteration
    n_c = timed["rollout_candidate"]["slots"] / max(timed["linearize"]["slots"] - B * args.steps, 1)
    f_iter = 100 * (F_LIN + F_PHI + F_COST + n_c * (F_BELL + F_ROLL + F_COST)) + 100 * 2 * (NX * NX + NX * NU)
    exec_total = sum(EXEC_FLOPS_PER_STEP[c] * 100 * timed[c]["slots"] for c in CLASSES)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * elapsed_s / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "bicycle car RK4 H=100 dt=0.01 contouring, ddp_linquad_reg (BASELINE configs[3], per-GPU slice)",
                   "batch_per_gpu": B, "global_batch": B * world, "solver_iterations": iters, "line_search_candidates": L, "scheduler": {0: "auto (lock-step)", 1: "lock-step", 2: "asynchronous"}[args.scheduler],
                   "parallelism": f"dp{world} (independent slices, final all-gather only)",
                   "l2": "working set (expansion+gains, >1 GB) is far larger than the 126 MB L2; no flush needed"},
        "e2e": {"value": e2e_tot.item() / e2e_s.item(), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps},
        "gpu_launches": n_launches,
        "clocks": clocks.summary(),
        "roofline": roofline,
        "per_step_ms": per_step_ms,
        "kernels": {c: {"ms": round(p["ms"], 3), "launches": p["launches"], "slots": p["slots"]} for c, p in prof.items()},
        "kernels_timed_region": {c: {"ms": round(p["ms"], 3), "launches": p["launches"], "slots": p["slots"]} for c, p in timed.items()},
        "solver": {"candidates_per_iteration": n_c, "algorithmic_flops_per_iteration": f_iter,
                   "algorithmic_tflops": f_iter * value / 1e12, "frac_of_fp64_peak": f_iter * value / peak_fp64,
                   "executed_tflops_whole_step": exec_total * world / elapsed_s / 1e12,
                   "executed_frac_of_fp64_peak": exec_total / elapsed_s / peak_fp64},
    }
    # ---- CPU baseline (rank 0, N=1 only): the oracle port on the host cores, bounded sample ----
    if world == 1 and not args.no_cpu_baseline:
        import multiprocessing as mp
        n_proc = min(os.cpu_count() or 1, args.cpu_procs)
        with mp.get_context("spawn").Pool(n_proc) as pool:
            c_its, c_wall = cpu_reference_step(pool, n_proc, 0)
        line["cpu_baseline"] = {"value": c_its / c_wall, "unit": UNIT, "cores": n_proc, "kind": "port",
                                "sample": f"{n_proc} problems of the same workload x 1 iteration, one process per core "
                                          f"(oracle/ilqc_oracle.py), {c_wall:.1f} s wall"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32768, help="problems per GPU")
    ap.add_argument("--iters", type=int, default=20, help="solver iterations per problem")
    ap.add_argument("--candidates", type=int, default=128, help="max line-search candidates per problem per round")
    ap.add_argument("--scheduler", type=int, default=0, help="0 auto, 1 lock-step iterations, 2 asynchronous")
    ap.add_argument("--cpu-procs", type=int, default=64)
    ap.add_argument("--ref-max-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
