#!/bin/bash
# 8-GPU validation of the driver's scaling launch: bench.py under torchrun (library NCCL, config-4/5 extras = the full configs 4 and 5)
set -x
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 8 --warmup 3 \
    2> gpurun_out/bench_r2_final_n8.err | tee gpurun_out/bench_r2_final_n8.json | cut -c1-300
tail -4 gpurun_out/bench_r2_final_n8.err
