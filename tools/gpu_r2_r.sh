#!/bin/bash
# attention kernel 6 (eight softmax warps per CTA): kernel tests, A/B
mkdir -p gpurun_out
P5_TEST_ATTN_IMPLS=6 timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k attention 2>&1 | tail -12 | tee gpurun_out/test_attn6.txt
timeout 600 python tools/ab_attention.py --iters 20 --out gpurun_out/ab_attention_r2r.json 2>&1 | grep "impl 31\|impl  6" | tee gpurun_out/ab_attention_r2r.txt
