#!/bin/bash
# GPU box: bench.py against library variants (ms per solve, kernel classes): bash tools/gpu_r2_variants.sh base pipe ...
mkdir -p gpurun_out
for t in "$@"; do
  LIBV=ilqc_b200/build/variants/libilqc_$t.so
  timeout 600 python tools/bench_with_lib.py $LIBV --steps 4 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/bench_var_$t.json 2> gpurun_out/bench_var_$t.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_var_$t.json"))
print("$t", round(d["ms_per_step"],2), {k:round(x["ms"],1) for k,x in d["kernels"].items()})
PY
done | tee gpurun_out/variants_bench.txt
