#!/bin/bash
# Attention feature A/B: kernel tests, isolated timings of every feature mask, in-step bench with masks 0 and 7.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv,noheader
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k attention 2>&1 | tail -5
timeout 300 python tools/ab_attention.py --iters 30 --out gpurun_out/ab_attention.json 2>&1 | tail -30
P5_ATTN_FEAT=0 timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_feat0.err | tee gpurun_out/bench_feat0.json | cut -c1-400
timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_feat7.err | tee gpurun_out/bench_feat7.json | cut -c1-400
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
