#!/bin/bash
# Round-2 evidence pass (1 GPU): full GPU suite, smoke, both bench arms, launch list of bench.py, ncu --set full of the
# projection GEMMs (-> profiles/ncu_traffic.json, tied to the build commit) and of the product attention kernel, sanitizers.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv,noheader
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | grep -v "^$" > gpurun_out/gputest_r2_final.txt; tail -5 gpurun_out/gputest_r2_final.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke_r2_final.txt
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 2> gpurun_out/bench_r2_final_ref.err | tee gpurun_out/bench_r2_final_ref.json | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 9 -c 4 -f -o gpurun_out/prof_gemm_r2 \
    python tools/profile_target.py > gpurun_out/prof_gemm_r2.log 2>&1
python tools/make_ncu_traffic.py gpurun_out/prof_gemm_r2.ncu-rep > gpurun_out/ncu_traffic_r2.json 2>&1 && cp profiles/ncu_traffic.json gpurun_out/ncu_traffic.json
tail -3 gpurun_out/prof_gemm_r2.log | cut -c1-200
timeout 900 python bench.py --steps 10 --warmup 3 2> gpurun_out/bench_r2_final.err | tee gpurun_out/bench_r2_final.json | cut -c1-400
tail -3 gpurun_out/bench_r2_final.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 513 -c 342 --csv --log-file gpurun_out/launches_bench_r2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/launches_bench_r2.log 2>&1
tail -2 gpurun_out/launches_bench_r2.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 2 -c 1 -f -o gpurun_out/prof_attn_r2 \
    python tools/profile_target.py > gpurun_out/prof_attn_r2.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_target.py 2>&1 | tail -4 | tee gpurun_out/sanitizer_memcheck_r2.txt
ls gpurun_out | tail -30
