#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/prof_quad.py 2>&1 | tail -1 | cut -c1-330
timeout 600 python -m pytest tests -m gpu -q -k "quad or random" 2>&1 | tail -2
