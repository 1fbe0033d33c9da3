#!/bin/bash
# N-GPU bench exactly as the driver launches it.  usage: bash tools/gpu_multi.sh N
N=${1:-2}
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/multi_gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"
cat gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench N=1 rc=$?"
cat gpurun_out/bench_n1.json
