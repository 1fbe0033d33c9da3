#!/bin/bash
set -x
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 2 --no-cpu-baseline"
run() {
  tag=$1; shift
  env "$@" timeout 600 $B --scheduler $SCHED > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$tag.json"))
print("$tag", "value %.0f" % d["value"], "e2e %.0f" % d["e2e"]["value"], "ms/step %.1f" % d["ms_per_step"], "n_c %.2f" % d["solver"]["candidates_per_iteration"], "launches", d["gpu_launches"])
PY
}
SCHED=2
for ws in 47360 56832 66304 75776 113664 151552; do run a2_w$ws ILQC_WAVE_SLOTS=$ws; done
SCHED=1
for ws in 56832 75776 151552; do run l2_w$ws ILQC_WAVE_SLOTS=$ws; done
