#!/bin/bash
# A multicast inside 8-CTA clusters (O / FFN-out projections): model + kernel tests, bench A/B on the debug library
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -x -q -m gpu -k "not attention" 2>&1 | tail -6 | tee gpurun_out/test_multicast.txt
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras --attn-impl 1 2> gpurun_out/bench_mc_on.err | tee gpurun_out/bench_mc_on.json | cut -c1-200
tail -2 gpurun_out/bench_mc_on.err
P5_GEMM_MULTICAST=0 timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras --attn-impl 1 2> gpurun_out/bench_mc_off.err | tee gpurun_out/bench_mc_off.json | cut -c1-200
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras --attn-impl 1 2> gpurun_out/bench_mc_on2.err | tee gpurun_out/bench_mc_on2.json | cut -c1-200
