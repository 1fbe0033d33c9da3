#!/bin/bash
# GPU box: second-order path timing + quad parity tests (+ ncu capture of the dynamics-Hessian kernel with NCU=1)
mkdir -p gpurun_out
timeout 600 python tools/prof_quad.py > gpurun_out/prof_quad_plain.log 2>&1; tail -1 gpurun_out/prof_quad_plain.log
timeout 900 python -m pytest tests -m gpu -q -k "quad or hess or expansion or random or cfg4 or config4" > gpurun_out/pytest_quad.log 2>&1; tail -3 gpurun_out/pytest_quad.log
if [ -n "$NCU" ]; then
ncu --set full --clock-control none --import-source on -k regex:"hess_kernel|backward_kernel" -s 6 -c 5 -o gpurun_out/prof_hess -f python tools/prof_quad.py > gpurun_out/ncu_hess.log 2>&1
echo "ncu rc=$?"
fi
ILQC_TRACE_ROUNDS=1 timeout 300 python tools/prof_mpc3.py > gpurun_out/prof_mpc3.log 2> gpurun_out/prof_mpc3_rounds.txt; tail -1 gpurun_out/prof_mpc3.log
