"""In-tree build of the CUDA library (and of the test-only host simulator).

    python -m ilqc_b200.build            # libilqc_b200.so  (nvcc, sm_100a; cross-compiles without a GPU)
    python -m ilqc_b200.build hostsim    # tests/hostsim/libilqc_hostsim.so (g++; test infrastructure only)
    python -m ilqc_b200.build all

One translation unit per model so the heavy templates compile in parallel.  Incremental: an object is
rebuilt only when a source it depends on is newer.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
HOSTSIM = os.path.join(ROOT, "tests", "hostsim")
CUDA_LIB = os.path.join(PKG, "libilqc_b200.so")
HOSTSIM_LIB = os.path.join(HOSTSIM, "libilqc_hostsim.so")
# model translation units: public ids 0-3 + the three-control-set (`rk4`) variants 5-7 (csrc/models.cuh)
MODEL_IDS = (0, 1, 2, 3, 5, 6, 7)

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]
GXX_FLAGS = ["-std=c++17", "-O1", "-fPIC", "-DILQC_HOSTSIM", "-ffp-contract=off", "-x", "c++"]


def _deps():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".inl", ".cu"))]
    files.append(os.path.join(ROOT, "include", "ilqc_b200.h"))
    return files


def _stale(target, extra=()):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(f) > t for f in list(_deps()) + list(extra))


def _run(cmd, log=None):
    p = subprocess.run(cmd, capture_output=True, text=True)
    if log:
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + p.stdout + p.stderr)
    if p.returncode != 0:
        raise RuntimeError("build failed: %s\n%s\n%s" % (" ".join(cmd), p.stdout[-4000:], p.stderr[-4000:]))
    return p.stdout + p.stderr


def build_cuda(verbose=True, force=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... -> ilqc_b200/libilqc_b200.so"""
    obj_dir = os.path.join(PKG, "build")
    os.makedirs(obj_dir, exist_ok=True)
    jobs = []
    for k in MODEL_IDS:
        jobs.append((os.path.join(obj_dir, f"model_{k}.o"),
                     ["nvcc", *NVCC_FLAGS, f"-DILQC_MODEL_INST={k}", "-c", os.path.join(CSRC, "model_inst.cu")]))
    jobs.append((os.path.join(obj_dir, "dense.o"), ["nvcc", *NVCC_FLAGS, "-c", os.path.join(CSRC, "dense.cu")]))
    jobs.append((os.path.join(obj_dir, "api.o"), ["nvcc", *NVCC_FLAGS, "-c", os.path.join(CSRC, "api.cu")]))

    def one(job):
        obj, cmd = job
        if force or _stale(obj):
            if verbose:
                print("[ilqc_b200.build] nvcc ->", os.path.basename(obj), flush=True)
            _run(cmd + ["-o", obj], log=obj + ".log")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        objs = list(ex.map(one, jobs))
    if force or _stale(CUDA_LIB, objs):
        _run(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", CUDA_LIB, *objs, "-lcudart"])
        if verbose:
            print("[ilqc_b200.build] linked", CUDA_LIB, flush=True)
    return CUDA_LIB


def build_hostsim(verbose=True, force=False):
    """g++ -DILQC_HOSTSIM ... -> tests/hostsim/libilqc_hostsim.so (CPU emulation, tests only)."""
    obj_dir = os.path.join(HOSTSIM, "build")
    os.makedirs(obj_dir, exist_ok=True)
    srcs = [os.path.join(HOSTSIM, f) for f in ("hostsim_api.cpp", "hostsim_model.cpp", "hostsim_dense.cpp")]
    jobs = []
    for k in MODEL_IDS:
        jobs.append((os.path.join(obj_dir, f"model_{k}.o"),
                     ["g++", *GXX_FLAGS, f"-DILQC_MODEL_INST={k}", "-c", srcs[1]]))
    jobs.append((os.path.join(obj_dir, "dense.o"), ["g++", *GXX_FLAGS, "-c", srcs[2]]))
    jobs.append((os.path.join(obj_dir, "api.o"), ["g++", *GXX_FLAGS, "-c", srcs[0]]))

    def one(job):
        obj, cmd = job
        if force or _stale(obj, srcs):
            if verbose:
                print("[ilqc_b200.build] g++ ->", os.path.basename(obj), flush=True)
            _run(cmd + ["-o", obj])
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        objs = list(ex.map(one, jobs))
    if force or _stale(HOSTSIM_LIB, objs):
        _run(["g++", "-shared", "-o", HOSTSIM_LIB, *objs, "-lm"])
        if verbose:
            print("[ilqc_b200.build] linked", HOSTSIM_LIB, flush=True)
    return HOSTSIM_LIB


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    what = args[0] if args else "cuda"
    if what in ("cuda", "all"):
        build_cuda(force="--force" in sys.argv)
    if what in ("hostsim", "all"):
        build_hostsim(force="--force" in sys.argv)
