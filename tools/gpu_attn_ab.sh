#!/bin/bash
# Attention feature A/B: kernel tests, isolated timings of feature masks, in-step bench, ncu of the default kernel.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv,noheader
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k attention 2>&1 | tail -5
timeout 300 python tools/ab_attention.py --iters 30 --out gpurun_out/ab_attention.json 2>&1 | tail -30
timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_feat_all.err | tee gpurun_out/bench_feat_all.json | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 2 -c 1 -f -o gpurun_out/prof_attn_tc \
    python tools/profile_target.py > gpurun_out/prof_attn_tc.log 2>&1
tail -2 gpurun_out/prof_attn_tc.log | cut -c1-300
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
