#!/bin/bash
# tests + bench + workloads of configs 4/5 (samples) + ncu of the attention kernel
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench.json
tail -3 gpurun_out/bench.err
timeout 300 python tools/profile_target.py --workload config4 --n 20000 --passes 2 2>&1 | tail -1 | tee gpurun_out/config4_20k.txt
timeout 300 python tools/profile_target.py --workload config5 --n 256 --passes 2 2>&1 | tail -1 | tee gpurun_out/config5_256.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python tools/profile_target.py --passes 2 > gpurun_out/launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 2 -c 1 -f -o gpurun_out/prof_attn_tc \
    python tools/profile_target.py > gpurun_out/prof_attn_tc.log 2>&1
tail -2 gpurun_out/prof_attn_tc.log
