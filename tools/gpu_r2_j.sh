#!/bin/bash
mkdir -p gpurun_out
(P5_NOMATH=1 timeout 300 python tools/ab_phase.py; P5_NOMATH=1 P5_ATTN_CTAS=1 timeout 300 python tools/ab_phase.py) > gpurun_out/ab_phase_nomath.txt 2>&1; cat gpurun_out/ab_phase_nomath.txt
