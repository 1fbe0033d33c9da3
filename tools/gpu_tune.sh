#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
B="python bench.py --steps 5 --warmup 2 --no-cpu-baseline"
run() {
  tag=$1; shift
  env "$@" timeout 600 $B --scheduler $SCHED > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$tag.json"))
print("$tag", "value %.0f" % d["value"], "e2e %.0f" % d["e2e"]["value"], "ms/step %.1f" % d["ms_per_step"], {k: round(v["ms"],1) for k,v in d["kernels"].items()}, "n_c %.2f" % d["solver"]["candidates_per_iteration"], "launches", d["gpu_launches"])
PY
}
SCHED=2
for ws in 94720 113664 132608 151552 189440; do run a_w$ws ILQC_WAVE_SLOTS=$ws; done
for ws in 94720 132608; do run a_prio_w$ws ILQC_WAVE_SLOTS=$ws ILQC_SIDE_PRIO=1; done
SCHED=1
for ws in 75776 94720; do run l_w$ws ILQC_WAVE_SLOTS=$ws; done
