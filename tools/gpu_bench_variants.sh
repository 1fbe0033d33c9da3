#!/bin/bash
# GPU box: short bench of the default library and of every variant of tools/build_variants.py
mkdir -p gpurun_out
run() {
  python - <<PY
import json
d=json.load(open("gpurun_out/bv.json")); k=d["kernels"]
print("$1", "value %.0f" % d["value"], "ms/step %.1f" % d["ms_per_step"], "lin %.1f bwd %.1f roll %.1f adj %.1f" % (k["linearize"]["ms"], k["backward"]["ms"], k["rollout_candidate"]["ms"], k["adjoint"]["ms"]))
PY
}
python bench.py --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/bv.json 2> gpurun_out/bv.err; run default
for lib in ilqc_b200/build/variants/libilqc_*.so; do
  python tools/bench_with_lib.py $lib --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/bv.json 2> gpurun_out/bv.err; run $lib
done
python bench.py --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/bv.json 2> gpurun_out/bv.err; run default_again
