#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "coop" > gpurun_out/pytest_coop.log 2>&1; tail -5 gpurun_out/pytest_coop.log | cut -c1-400
timeout 900 python tools/bench_coop.py > gpurun_out/bench_coop.txt 2> gpurun_out/bench_coop.err; cat gpurun_out/bench_coop.txt; tail -3 gpurun_out/bench_coop.err
