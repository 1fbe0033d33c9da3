"""Rewrite profiles/kernel_counters.json (read by bench.py for the executed-flop / DRAM-traffic lines of its roofline
object) from an `ncu --set full` capture of bench.py's kernels, and print the per-kernel summary kept under profiles/.

    python tools/ncu_counters.py gpurun_out/prof_hot.ncu-rep [--horizon 100] [--write] [--note "..."]

Per class the launch with the largest grid is used (the full first line-search round / the full-batch expansion):
exec flops per (slot, time step) = (dadd + dmul + 2 dfma thread instructions) / (threads x horizon), DRAM bytes per slot =
(dram__bytes_read.sum + dram__bytes_write.sum) / threads."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLASSES = (("forward_kernel", ", 1>", "linearize"), ("forward_kernel", ", 0>", "rollout_cost"), ("backward_kernel", "", "backward"),
           ("rollout_kernel", "", "rollout_candidate"), ("adjoint_kernel", "", "adjoint"), ("hess_kernel", "", "hess_tensor"))


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return float("nan")


def main():
    rep = sys.argv[1]
    H = int(sys.argv[sys.argv.index("--horizon") + 1]) if "--horizon" in sys.argv else 100
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, rows = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def scaled(r, name):   # value in base units (ncu prints Gbyte / Mbyte / ms ...)
        v, u = num(r[col[name]]), units[col[name]]
        mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u, 1.0)
        return v * mult
    best = {}
    lines = []
    for r in rows:
        name = r[col["Kernel Name"]]
        threads = num(r[col["launch__grid_size"]]) * num(r[col["launch__block_size"]])
        # (capture with --metrics smsp__sass_thread_inst_executed_op_{dadd,dmul,dfma}_pred_on.sum on top of --set full)
        fl = sum(num(r[col[f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum"]]) * w
                 for op, w in (("dadd", 1), ("dmul", 1), ("dfma", 2))) if "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum" in col else float("nan")
        dram = scaled(r, "dram__bytes_read.sum") + scaled(r, "dram__bytes_write.sum")
        t = scaled(r, "gpu__time_duration.sum")
        stall = sorted(((num(r[i]), h.split("issue_stalled_")[1].split("_per")[0]) for h, i in col.items()
                        if "smsp__average_warp" in h and "issue_stalled" in h and "not_issued" not in h and r[i]), reverse=True)[:4]
        lines.append(f"{name[:70]} | threads {int(threads)} | {t * 1e3:.3f} ms | regs {r[col['launch__registers_per_thread']]} | "
                     f"{fl / t / 1e12:.1f} TFLOP/s executed | {fl / threads / H:.1f} flops/slot/step | DRAM {dram / 1e9:.3f} GB = "
                     f"{dram / threads / 1e3:.1f} KB/slot = {dram / t / 1e12:.2f} TB/s | fp64 pipe "
                     f"{r[col['sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active']]} % | issue "
                     f"{r[col['sm__issue_active.avg.pct_of_peak_sustained_elapsed']]} % | warps_active {r[col['sm__warps_active.avg.pct_of_peak_sustained_active']]} % | "
                     "stalls " + ", ".join(f"{n}={v:.2f}" for v, n in stall))
        for key, tag, cls in CLASSES:
            if key in name and tag in name and (cls not in best or threads > best[cls][0]):
                best[cls] = (threads, fl / threads / H, dram / threads)
    print("\n".join(lines))
    if "--write" in sys.argv:
        path = os.path.join(ROOT, "profiles", "kernel_counters.json")
        old = json.load(open(path))
        commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
        old["commit"] = commit
        old["captured"] = (sys.argv[sys.argv.index("--note") + 1] if "--note" in sys.argv else "") + " (" + os.path.basename(rep) + ")"
        for cls, (threads, f, d) in best.items():
            if cls in old["classes"]:
                old["classes"][cls] = {"exec_flops_per_slot_step": round(f, 1), "dram_bytes_per_slot": round(d), "threads_of_capture": int(threads)}
        json.dump(old, open(path, "w"), indent=1)
        print("wrote", path, "at commit", commit)


if __name__ == "__main__":
    main()
