#!/bin/bash
# GPU box: parity tests, then the bench for several speculation capacities (ILQC_WAVE_SLOTS = candidate slots
# evaluated per line-search round) of the lock-step scheduler.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline"
run() {
  tag=$1; shift
  env "$@" timeout 600 $B > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$tag.json"))
k=d["kernels"]
print("$tag", "value %.0f" % d["value"], "e2e %.0f" % d["e2e"]["value"], "ms/step %.1f" % d["ms_per_step"], "n_c %.2f" % d["solver"]["candidates_per_iteration"], "launches", d["gpu_launches"], "lin %.1f bwd %.1f roll %.1f" % (k["linearize"]["ms"], k["backward"]["ms"], k["rollout_candidate"]["ms"]))
PY
}
for ws in 47360 56832 66304 75776 94720 113664 151552; do run w$ws ILQC_WAVE_SLOTS=$ws; done
