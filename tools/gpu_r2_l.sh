#!/bin/bash
mkdir -p gpurun_out
P5_KERNEL4=1 timeout 300 python tools/ab_phase.py 2>&1 | tee gpurun_out/ab_phase_k4.txt
