#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 4 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"
head -c 330 gpurun_out/bench_n$N.json; echo
