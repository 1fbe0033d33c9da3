#!/bin/bash
mkdir -p gpurun_out
for t in base tpf base tpf; do
echo "== $t"; if [ $t = base ]; then L=ilqc_b200/libilqc_b200.so; else L=ilqc_b200/build/variants/libilqc_$t.so; fi
ILQC_LIB=$L timeout 600 python tools/prof_quad.py 2>&1 | tail -1 | cut -c1-300
done
