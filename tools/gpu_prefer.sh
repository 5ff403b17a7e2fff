#!/bin/bash
set -x
mkdir -p gpurun_out
P5_GEMM_CLUSTER=8 P5_GEMM_PREFER=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k gemm 2>&1 | tail -4
SWEEP_PREFER_ONLY=1 timeout 900 python tools/sweep_gemm_traffic.py ffn_out o ffn_in 2>&1 | tee gpurun_out/gemm_prefer_sweep.txt | tail -30
timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2> /dev/null | python -c "import sys,json; b=json.loads(sys.stdin.read()); print('default', b['ms_per_step'], b['roofline']['achieved'], b['clocks'])"
P5_GEMM_CLUSTER=8 P5_GEMM_PREFER=1 timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2> /dev/null | python -c "import sys,json; b=json.loads(sys.stdin.read()); print('prefer8', b['ms_per_step'], b['roofline']['achieved'], b['clocks'])"
timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2> /dev/null | python -c "import sys,json; b=json.loads(sys.stdin.read()); print('default', b['ms_per_step'], b['roofline']['achieved'], b['clocks'])"
