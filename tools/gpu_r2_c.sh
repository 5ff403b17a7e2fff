#!/bin/bash
# attention kernel 3 (packed-pair math): tests, A/B, bench with it, ncu
set -x
mkdir -p gpurun_out
P5_TEST_ATTN_IMPLS=3 timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention" 2>&1 | tail -8
timeout 600 python tools/ab_attention.py --iters 30 --out gpurun_out/ab_attention_r2c.json 2>&1 | tail -30
timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras --attn-impl 3 2> gpurun_out/bench_r2c_impl3.err | tee gpurun_out/bench_r2c_impl3.json | cut -c1-300
tail -3 gpurun_out/bench_r2c_impl3.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 2 -c 1 -f -o gpurun_out/prof_attn_impl3 \
      python tools/attn_target.py 3 config2 > gpurun_out/prof_attn_impl3.log 2>&1
tail -2 gpurun_out/prof_attn_impl3.log
