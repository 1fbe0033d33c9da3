#!/bin/bash
# GPU box: kernel timings of every library variant under ilqc_b200/build/variants (tools/build_variants.py)
mkdir -p gpurun_out
timeout 1200 python tools/tune_kernels.py > gpurun_out/tuning_variants.txt 2> gpurun_out/tuning_variants.err; echo "tune rc=$?"
cat gpurun_out/tuning_variants.txt
