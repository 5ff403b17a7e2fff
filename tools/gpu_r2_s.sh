#!/bin/bash
# fused RMSNorm in the residual-add GEMM epilogue: model + norm tests, bench with and without
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -x -q -m gpu -k "not attention" 2>&1 | tail -6 | tee gpurun_out/test_fused_norm.txt
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras --fuse-norm 2> gpurun_out/bench_fused.err | tee gpurun_out/bench_fused.json | cut -c1-250
tail -2 gpurun_out/bench_fused.err
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras 2> gpurun_out/bench_unfused.err | tee gpurun_out/bench_unfused.json | cut -c1-250
