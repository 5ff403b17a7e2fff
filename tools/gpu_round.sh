#!/bin/bash
# One gpurun call: GPU tests, bench (both arms), ncu launch list + full captures of the top kernels.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
nproc; lscpu | grep "Model name"
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 600 python bench.py --steps 10 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench.json
tail -5 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2> gpurun_out/bench_ref.err | tee gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python tools/profile_target.py --passes 2 > gpurun_out/launches.log 2>&1
tail -2 gpurun_out/launches.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 9 -c 4 -f -o gpurun_out/prof_gemm \
    python tools/profile_target.py > gpurun_out/prof_gemm.log 2>&1
tail -2 gpurun_out/prof_gemm.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention -s 2 -c 1 -f -o gpurun_out/prof_attn \
    python tools/profile_target.py > gpurun_out/prof_attn.log 2>&1
tail -2 gpurun_out/prof_attn.log
ls -la gpurun_out
