#!/bin/bash
# attention kernel 4: correctness (kernel-level tests restricted to impl 4), A/B timing, phase counters
mkdir -p gpurun_out
P5_TEST_ATTN_IMPLS=4 timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k attention 2>&1 | tail -15 | tee gpurun_out/test_attn4.txt
timeout 600 python tools/ab_attention.py --iters 20 --out gpurun_out/ab_attention_r2k.json 2>&1 | grep "impl 31\|impl  4" | tee gpurun_out/ab_attention_r2k.txt
P5_KERNEL4=1 timeout 300 python tools/ab_phase.py 2>&1 | tee gpurun_out/ab_phase_k4.txt
