#!/bin/bash
# GPU box: randomised-configuration soak on the CUDA engine + 2-GPU bench with the final kernels
mkdir -p gpurun_out
timeout 1200 python tools/soak_random_configs.py 16 96 2>&1 | grep -v -i warn | tail -3 | tee gpurun_out/soak_gpu.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?"
head -c 400 gpurun_out/bench_n2.json; echo
