set -x
mkdir -p gpurun_out
PCMD="python bench.py --steps 1 --warmup 1 --iters 2 --no-cpu-baseline"
$PCMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"rollout_kernel|backward_kernel|forward_kernel" -s 4 -c 6 -o gpurun_out/prof_hot_fm -f $PCMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
