#!/bin/bash
# Round 2: phase counters of the attention softmax warps (1 GPU), one and two CTAs per SM.
mkdir -p gpurun_out
(timeout 300 python tools/ab_phase.py; P5_ATTN_CTAS=1 timeout 300 python tools/ab_phase.py) > gpurun_out/ab_phase.txt 2>&1; cat gpurun_out/ab_phase.txt
