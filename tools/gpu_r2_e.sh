#!/bin/bash
# Round 2 validation on a 2-GPU box: full GPU suite (incl. --procs 2 and the in-process 2-device test), smoke, bench at
# N=1 (full line), the reference arm, bench at N=2 (library NCCL all-gather + config-4/5 extras).
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv,noheader
timeout 1800 python -m pytest tests -x -q -m gpu -s 2>&1 | grep -v "^$" > gpurun_out/gputest_r2e.txt; tail -40 gpurun_out/gputest_r2e.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 2> gpurun_out/bench_r2e_n1.err | tee gpurun_out/bench_r2e_n1.json | cut -c1-600
tail -3 gpurun_out/bench_r2e_n1.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 2> gpurun_out/bench_r2e_ref.err | tee gpurun_out/bench_r2e_ref.json | cut -c1-600
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 \
    2> gpurun_out/bench_r2e_n2.err | tee gpurun_out/bench_r2e_n2.json | cut -c1-600
tail -5 gpurun_out/bench_r2e_n2.err
