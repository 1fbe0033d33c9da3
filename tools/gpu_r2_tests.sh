#!/bin/bash
# GPU box (round 2): all gpu tests (no -x), teacher-forced rows above 1e-10 dumped for tests/golden/make_sensitivity.py
mkdir -p gpurun_out
export ILQC_DUMP_ROWS=gpurun_out/tf_rows
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/pytest_gpu.log | cut -c1-400 | tail -40
timeout 600 python tools/prof_quad.py > gpurun_out/prof_quad_plain.log 2>&1; tail -2 gpurun_out/prof_quad_plain.log
