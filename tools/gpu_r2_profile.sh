#!/bin/bash
# GPU box (round 2 evidence set), two calls because a call may bring back at most 64 MiB:
#   bash tools/gpu_r2_profile.sh hot   gpu tests, full bench line, ncu launch list of the bench command, full capture of the hot kernels
#   bash tools/gpu_r2_profile.sh quad  second-order path: timing + full capture of the tensor kernel and the contracting backward pass
mkdir -p gpurun_out
OPS=smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum
if [ "$1" = "hot" ]; then
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"
head -c 600 gpurun_out/bench_full.json; echo
PCMD="python bench.py --steps 1 --warmup 1 --iters 2 --no-cpu-baseline --no-other-configs"
$PCMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $PCMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
$PCMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --metrics $OPS --clock-control none --import-source on -k regex:"rollout_kernel|backward_kernel|forward_kernel|adjoint_kernel" -s 5 -c 6 -o gpurun_out/prof_hot -f $PCMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
else
python tools/prof_quad.py > gpurun_out/prof_quad_plain.log 2>&1 && \
ncu --set full --metrics $OPS --clock-control none --import-source on -k regex:"hess_kernel|backward_kernel" -s 8 -c 4 -o gpurun_out/prof_quad -f python tools/prof_quad.py > gpurun_out/ncu_quad.log 2>&1
echo "ncu quad rc=$?"; tail -1 gpurun_out/prof_quad_plain.log
timeout 600 python -m pytest tests -m gpu -q -k "quad or random or cfg4 or coop" > gpurun_out/pytest_quad.log 2>&1; tail -2 gpurun_out/pytest_quad.log
fi
ls -la gpurun_out | tail -12
