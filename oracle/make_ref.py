"""TEST / BASELINE INFRASTRUCTURE ONLY -- recipe that populates `oracle/_ref/` with the UNMODIFIED reference.

    python oracle/make_ref.py            # in the build container (needs /root/reference)

vroulet/ilqc is pure Python: there is nothing to compile, its "build" is a verbatim copy of the package directories the
hot path imports (algorithms/, envs/ with its track / model JSON files, model_predictive_control/, utils_pipeline/).  `oracle/_ref/` is
git-ignored (reference sources are never committed to this repository) but NOT gpurun-ignored, so it travels to the GPU
box like a built `.so`, where `bench.py --impl reference` and `bench.py`'s `cpu_baseline` leg time the reference's own
`run_min_algo` on the host cores (`cpu_baseline.kind = "reference"`).  When the directory is absent those legs fall back
to the oracle port (`kind = "port"`).  A MANIFEST with the sha256 of every copied file is written next to them;
`load_ref()` re-checks it before importing, so a modified copy is refused.

Nothing under ilqc_b200/ imports this module.
"""
import hashlib
import json
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
REF_DST = os.path.join(HERE, "_ref")
PACKAGES = ("algorithms", "envs", "model_predictive_control", "utils_pipeline")
KEEP_EXT = (".py", ".json", ".csv")


def _sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def make_ref(verbose=True):
    if not os.path.isdir(REF_SRC):
        raise SystemExit(f"{REF_SRC} not present: oracle/_ref can only be populated in the build container")
    if os.path.isdir(REF_DST):
        shutil.rmtree(REF_DST)
    manifest = {}
    for pkg in PACKAGES:
        for root, _, files in os.walk(os.path.join(REF_SRC, pkg)):
            for f in files:
                if not f.endswith(KEEP_EXT):
                    continue
                src = os.path.join(root, f)
                rel = os.path.relpath(src, REF_SRC)
                dst = os.path.join(REF_DST, rel)
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copyfile(src, dst)
                manifest[rel] = _sha(dst)
    json.dump(manifest, open(os.path.join(REF_DST, "MANIFEST.json"), "w"), indent=0, sort_keys=True)
    if verbose:
        print(f"[oracle/make_ref] copied {len(manifest)} files of the unmodified reference into {REF_DST}")
    return REF_DST


def ref_available():
    return os.path.exists(os.path.join(REF_DST, "MANIFEST.json"))


def load_ref():
    """Import the unmodified reference from oracle/_ref (plotting modules stubbed: matplotlib / seaborn are not installed
    and carry no arithmetic).  Returns a namespace with run_min_algo, make_env, mpc_step."""
    manifest = json.load(open(os.path.join(REF_DST, "MANIFEST.json")))
    for rel, sha in manifest.items():
        if _sha(os.path.join(REF_DST, rel)) != sha:
            raise RuntimeError(f"oracle/_ref/{rel} differs from the reference it was copied from")
    for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.collections", "matplotlib.patches", "seaborn"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib.collections"].PatchCollection = object
    sys.modules["matplotlib.collections"].LineCollection = object
    sys.modules["matplotlib.patches"].Polygon = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REF_DST not in sys.path:
        sys.path.insert(0, REF_DST)
    import torch
    torch.set_default_dtype(torch.float64)      # what the reference's callers do (tests/test_newton.py:12)
    import algorithms.run_min_algo as rma
    from envs.choose_env import make_env
    from model_predictive_control.mpc_pipeline import mpc_step
    rma.print_info_step = lambda *a, **k: None   # silence the per-iteration table printer (no arithmetic)
    return types.SimpleNamespace(run_min_algo=rma.run_min_algo, make_env=make_env, mpc_step=mpc_step, rma=rma)


if __name__ == "__main__":
    make_ref()
