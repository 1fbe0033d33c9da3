#!/bin/bash
# 2-GPU sanity of the bench contract (torchrun launch as the driver does it) + the reference arm
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?"
head -c 700 gpurun_out/bench_n2.json; echo; tail -3 gpurun_out/bench_n2.err
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; head -c 900 gpurun_out/bench_ref.json; echo
