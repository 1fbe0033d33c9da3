#!/bin/bash
# quick GPU check: kernel timings (default lib only) + bench
set -x
mkdir -p gpurun_out
ILQC_TUNE_DEFAULT_ONLY=1 timeout 600 python tools/tune_kernels.py > gpurun_out/tuning_quick.txt 2> gpurun_out/tuning_quick.err
cat gpurun_out/tuning_quick.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_quick.json"))
print("value %.0f" % d["value"], "e2e %.0f" % d["e2e"]["value"], "ms/step %.1f" % d["ms_per_step"], "launches", d["gpu_launches"])
print("isolated", {k: (round(v["ms"],1), v["launches"]) for k,v in d["kernels"].items()})
print("roofline", {k: d["roofline"][k] for k in ("kernel","achieved","frac","ms_per_launch_avg","share_of_kernel_time")})
PY
tail -3 gpurun_out/bench_quick.err
