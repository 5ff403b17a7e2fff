#!/bin/bash
# Round 2, first GPU pass: new attention kernel (impl 2) tests + A/B, C oracle speed on the box, full GPU suite with the
# full-size fixtures, bench with both attention kernels.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv,noheader
nproc; grep -m1 "model name" /proc/cpuinfo
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention" 2>&1 | tail -15
timeout 600 python tools/ab_attention.py --iters 30 --out gpurun_out/ab_attention_r2a.json 2>&1 | tail -30
timeout 600 python - <<'PY' 2>&1 | tail -5
import os, sys, time
sys.path.insert(0, '.')
from oracle import prostt5_oracle_c as OC
from unicore_b200 import prostt5_spec as spec, synth
d = synth.model_dir('/tmp/p5_full_seed1', spec.FULL, seed=1)
oc = OC.load_gguf_model(d + '/' + spec.WEIGHT_FILE)
aa, off = spec.synthetic_proteome("config2")
oc.predict(aa[:350].tobytes())
t = time.time(); n = 6
for i in range(n): oc.predict(aa[350*i:350*(i+1)].tobytes())
dt = time.time() - t
print("C oracle on the box: %d threads, avx512 %s, %.1f residues/s, %.0f GFLOP/s" % (oc.threads, oc.uses_avx512, 350*n/dt, n*spec.FULL.flops_per_seq(350)/dt/1e9))
PY
timeout 1500 python -m pytest tests -x -q -m gpu -s 2>&1 | grep -v "^$" | tail -60 > gpurun_out/gputest_r2a.txt; tail -25 gpurun_out/gputest_r2a.txt
timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_r2a_impl1.err | tee gpurun_out/bench_r2a_impl1.json | cut -c1-400
timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --attn-impl 2 2> gpurun_out/bench_r2a_impl2.err | tee gpurun_out/bench_r2a_impl2.json | cut -c1-400
