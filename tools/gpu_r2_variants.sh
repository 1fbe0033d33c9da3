#!/bin/bash
mkdir -p gpurun_out
for t in hg2 hg4; do
  echo "== $t"; ILQC_LIB=ilqc_b200/build/variants/libilqc_$t.so timeout 600 python tools/prof_quad.py 2>&1 | tail -1 | cut -c1-300
done | tee gpurun_out/variants_hess.txt
