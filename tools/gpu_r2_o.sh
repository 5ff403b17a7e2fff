#!/bin/bash
# one-pass softmax in the product kernel (mask 31): kernel tests, A/B, phase counters
mkdir -p gpurun_out
P5_TEST_ATTN_IMPLS=1,47 timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k attention 2>&1 | tail -8 | tee gpurun_out/test_attn_onepass.txt
timeout 600 python tools/ab_attention.py --iters 20 --out gpurun_out/ab_attention_r2o.json 2>&1 | grep "impl 31\|impl 47\|impl  4" | tee gpurun_out/ab_attention_r2o.txt
timeout 300 python tools/ab_phase.py 2>&1 | tee gpurun_out/ab_phase_onepass.txt
