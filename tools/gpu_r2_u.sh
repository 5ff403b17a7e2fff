#!/bin/bash
# 2-GPU validation of the final state: the multi-device tests, bench at N=2 (library NCCL all-gather, config-4/5 extras)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_host.py tests/test_gpu_model.py -x -q -m gpu -k "two_ranks or multi_device or comm or fused" 2>&1 | tail -4 | tee gpurun_out/gputest_r2_2gpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 \
    2> gpurun_out/bench_r2_final_n2.err | tee gpurun_out/bench_r2_final_n2.json | cut -c1-300
tail -3 gpurun_out/bench_r2_final_n2.err
