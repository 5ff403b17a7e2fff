#!/bin/bash
# Full evidence pass: GPU tests, both bench arms, ncu launch list OF bench.py, ncu --set full of the two top kernels,
# attention A/B table, compute-sanitizer.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv,noheader
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2> gpurun_out/bench_ref.err | tee gpurun_out/bench_ref.json | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench.json
tail -3 gpurun_out/bench.err
timeout 300 python tools/ab_attention.py --iters 30 --out gpurun_out/ab_attention.json 2>&1 | tail -24
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 513 -c 342 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
tail -2 gpurun_out/launches_bench.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 9 -c 4 -f -o gpurun_out/prof_gemm \
    python tools/profile_target.py > gpurun_out/prof_gemm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 2 -c 1 -f -o gpurun_out/prof_attn_tc \
    python tools/profile_target.py > gpurun_out/prof_attn_tc.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_target.py 2>&1 | tail -4 | tee gpurun_out/sanitizer_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_target.py 2>&1 | tail -4 | tee gpurun_out/sanitizer_racecheck.txt
ls gpurun_out
