#!/bin/bash
# Final pass of a round: GPU tests, bench, ncu --set full of one layer's GEMMs, launch list of bench.py.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv,noheader
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-200
P5_GEMM_CLUSTER=2 timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2> /dev/null | python -c "import sys,json; b=json.loads(sys.stdin.read()); print('pairs-only', b['ms_per_step'], b['roofline']['achieved'], b['clocks'])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 9 -c 4 -f -o gpurun_out/prof_gemm \
    python tools/profile_target.py > gpurun_out/prof_gemm.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 513 -c 342 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
timeout 600 python tools/profile_target.py --workload config4 2>&1 | tail -1 | tee gpurun_out/config4_full.txt
