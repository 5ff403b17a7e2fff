#!/bin/bash
mkdir -p gpurun_out
P5_TEST_ATTN_IMPLS=5 timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k attention 2>&1 | tail -3
timeout 600 python tools/ab_attention.py --iters 20 2>&1 | grep "impl 31\|impl  5" | tee gpurun_out/ab_attention_r2t.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_tc5 -s 2 -c 1 -f -o gpurun_out/prof_attn_k5_long \
      python tools/attn_target.py 5 long > gpurun_out/prof_attn_k5_long.log 2>&1
tail -2 gpurun_out/prof_attn_k5_long.log
