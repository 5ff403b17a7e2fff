#!/bin/bash
set -x
mkdir -p gpurun_out
P5_GEMM_CLUSTER=8 timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k gemm 2>&1 | tail -4
P5_GEMM_CLUSTER=4 timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k gemm 2>&1 | tail -4
SWEEP_CLUSTER_ONLY=1 timeout 900 python tools/sweep_gemm_traffic.py ffn_out o ffn_in 2>&1 | tee gpurun_out/gemm_cluster_sweep.txt | tail -30
