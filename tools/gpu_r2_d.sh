#!/bin/bash
set -x
mkdir -p gpurun_out
P5_TEST_ATTN_IMPLS=3 timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention" 2>&1 | tail -4
timeout 600 python tools/ab_attention.py --iters 30 --out gpurun_out/ab_attention_r2d.json 2>&1 | grep -E "impl 31|impl  3"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 2 -c 1 -f -o gpurun_out/prof_attn_impl3b \
      python tools/attn_target.py 3 config2 > gpurun_out/prof_attn_impl3b.log 2>&1
tail -1 gpurun_out/prof_attn_impl3b.log
