#!/bin/bash
mkdir -p gpurun_out
P5_TEST_ATTN_IMPLS=16415 timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention_matches or many_items or peaked or neighbour" 2>&1 | tail -3
timeout 600 python tools/ab_attention.py --iters 30 2>&1 | grep "impl    31\|impl 16415" | tee gpurun_out/ab_attention_r2z4.txt
