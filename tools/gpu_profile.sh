#!/bin/bash
# GPU box: gpu tests, full bench, ncu launch list + full captures of the hot kernels.  Outputs: gpurun_out/.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"
cat gpurun_out/bench_full.json
PCMD="python bench.py --steps 1 --warmup 1 --iters 2 --no-cpu-baseline"
$PCMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $PCMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
$PCMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"rollout_kernel|backward_kernel|forward_kernel" -s 4 -c 6 -o gpurun_out/prof_hot -f $PCMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out
