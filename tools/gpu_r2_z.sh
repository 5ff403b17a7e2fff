#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/ab_attention.py --iters 30 2>&1 | grep "impl   31\|impl 1055" | tee gpurun_out/ab_attention_r2z.txt
