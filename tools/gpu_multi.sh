#!/bin/bash
# multi-GPU checks: in-process 2-device test, torchrun bench at N=$1
N=${1:-2}
set -x
nvidia-smi --query-gpu=index,name --format=csv
timeout 600 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "multi_device or config2" 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 6 --warmup 3 2> gpurun_out/bench_n$N.err | tee gpurun_out/bench_n$N.json
tail -5 gpurun_out/bench_n$N.err
