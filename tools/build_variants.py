"""Build tuning variants of libilqc_b200.so (only the bicycle-model translation unit differs).

    python tools/build_variants.py            # -> ilqc_b200/build/variants/libilqc_<tag>.so
Each variant = a set of -D flags for model_inst.cu (launch bounds / block size).  Used by tools/tune_kernels.py.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ilqc_b200 import build as b  # noqa: E402

VARIANTS = {
    "tpf": ["-DILQC_TENSOR_PREFETCH=1"],
}


def main():
    b.build_cuda(verbose=True)
    vdir = os.path.join(b.PKG, "build", "variants")
    os.makedirs(vdir, exist_ok=True)
    objs = [os.path.join(b.PKG, "build", f) for f in ("model_0.o", "model_1.o", "model_2.o", "model_5.o", "model_6.o", "model_7.o", "dense.o", "api.o")]

    def one(item):
        tag, flags = item
        obj = os.path.join(vdir, f"model_3_{tag}.o")
        lib = os.path.join(vdir, f"libilqc_{tag}.so")
        if not os.path.exists(lib) or b._stale(lib):
            cmd = ["nvcc", *b.NVCC_FLAGS, "-DILQC_MODEL_INST=3", *flags, "-c", os.path.join(b.CSRC, "model_inst.cu"), "-o", obj]
            p = subprocess.run(cmd, capture_output=True, text=True)
            open(obj + ".log", "w").write(p.stdout + p.stderr)
            assert p.returncode == 0, p.stderr[-2000:]
            subprocess.check_call(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib, *objs, obj, "-lcudart"])
        regs = [l for l in open(obj + ".log") if "registers" in l or "spill" in l]
        print(tag, "built", flush=True)
        return tag

    with ThreadPoolExecutor(max_workers=8) as ex:
        list(ex.map(one, VARIANTS.items()))


if __name__ == "__main__":
    main()
