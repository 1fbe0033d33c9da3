#!/bin/bash
# Scaling run on one box exactly as the driver launches it: N = 1, 2, 4, 8 (as many as are visible).
set -x
mkdir -p gpurun_out
NG=$(nvidia-smi --query-gpu=index --format=csv,noheader | wc -l)
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/multi_gpus.txt
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err; echo "bench N=1 rc=$?"
for N in 2 4 8; do
  if [ $N -le $NG ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err; echo "bench N=$N rc=$?"
    tail -3 gpurun_out/scale_n$N.err
  fi
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/scale_n*.json")):
    try:
        d=json.load(open(f)); print(f, d["n_gpus"], "value %.0f"%d["value"], "e2e %.0f"%d["e2e"]["value"], "ms/step %.1f"%d["ms_per_step"])
    except Exception as e: print(f, e)
PY
