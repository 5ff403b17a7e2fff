#!/bin/bash
# Round 2: pipe-rate microbenchmarks and timing ablations of the attention kernel (1 GPU).
mkdir -p gpurun_out
timeout 300 build/microbench > gpurun_out/microbench.txt 2>&1; tail -70 gpurun_out/microbench.txt
(timeout 300 python tools/ab_ablate.py; P5_ATTN_CTAS=1 timeout 300 python tools/ab_ablate.py) > gpurun_out/ab_ablate.txt 2>&1; cat gpurun_out/ab_ablate.txt
