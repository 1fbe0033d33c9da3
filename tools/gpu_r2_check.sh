#!/bin/bash
# GPU box (round 2): all gpu tests (no -x: report every failure), smoke, full bench line, second-order bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/pytest_gpu.log | tail -30
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"
cat gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
timeout 600 python tools/prof_quad.py > gpurun_out/prof_quad_plain.log 2>&1; cat gpurun_out/prof_quad_plain.log | tail
