#!/bin/bash
# GPU box: ncu capture of the second-order path (dynamics Hessian tensor kernel + tensor-contracting backward pass)
mkdir -p gpurun_out
python tools/prof_quad.py > gpurun_out/prof_quad_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"hess_kernel|backward_kernel" -s 6 -c 5 -o gpurun_out/prof_quad -f python tools/prof_quad.py > gpurun_out/ncu_quad.log 2>&1
echo "ncu rc=$?"
