#!/bin/bash
# GPU box: A/B of library variants (tools/build_variants.py): fastmath self-test, kernel timings, short bench per variant.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fastmath.py tests/test_gpu_parity.py -m gpu -x -q -s > gpurun_out/pytest_ab.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_ab.log
timeout 900 python tools/tune_kernels.py > gpurun_out/tuning_ab.txt 2> gpurun_out/tuning_ab.err; echo "tune rc=$?"
cat gpurun_out/tuning_ab.txt
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_ab.json"))
print("value %.0f" % d["value"], "e2e %.0f" % d["e2e"]["value"], "ms/step %.1f" % d["ms_per_step"], "n_c %.2f" % d["solver"]["candidates_per_iteration"], d["kernels"])
PY
