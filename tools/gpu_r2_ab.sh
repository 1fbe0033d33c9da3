#!/bin/bash
# GPU box (round 2): all gpu tests, then bench A/B of the small-round operand feed (ILQC_FEED_SMALL)
mkdir -p gpurun_out
export ILQC_DUMP_ROWS=gpurun_out/tf_rows
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/pytest_gpu.log | cut -c1-500 | tail -30
unset ILQC_DUMP_ROWS
for v in 0 16384 49152; do
  ILQC_FEED_SMALL=$v timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_feed_$v.json 2> gpurun_out/bench_feed_$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_feed_$v.json"))
print("FEED_SMALL=$v", d["ms_per_step"], d["value"], {k:(round(x["ms"],1)) for k,x in d["kernels"].items()}, "cfg5", d.get("other_configs",{}).get("cfg5_mpc_complex_track_4096_episodes_40_windows",{}).get("ms_per_window_solve"), "cfg4q", d.get("other_configs",{}).get("cfg4_bicycle_H100_B32768_ddp_quad_reg",{}).get("ms_per_solve"))
PY
done
