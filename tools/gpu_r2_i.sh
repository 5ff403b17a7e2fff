#!/bin/bash
mkdir -p gpurun_out
(timeout 300 python tools/ab_ablate.py; P5_ATTN_CTAS=1 timeout 300 python tools/ab_ablate.py) 2>&1 | grep "product\|at all" | tee gpurun_out/ab_ablate_nomath.txt
