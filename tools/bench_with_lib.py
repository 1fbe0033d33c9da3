"""Tuning aid: run bench.py against a library variant of tools/build_variants.py.
    python tools/bench_with_lib.py ilqc_b200/build/variants/libilqc_<tag>.so [bench.py arguments]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ilqc_b200 import _lib  # noqa: E402

_lib.CUDA_LIB_PATH = os.path.abspath(sys.argv[1])
sys.argv = ["bench.py"] + sys.argv[2:]
import bench  # noqa: E402

bench.main()
