#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
python -c "from unicore_b200 import synth, prostt5_spec as s; synth.model_dir('/tmp/p5_full_seed1', s.FULL, seed=1)"
for N in ${GPUS:-8 4}; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
    bench.py --gpus $N --steps 6 --warmup 3 2> gpurun_out/bench_n$N.err | tee gpurun_out/bench_n$N.json | cut -c1-400
tail -3 gpurun_out/bench_n$N.err
done
# in-process 8-device run of the C++ CLI on a config-4 sample (20k sequences)
python - <<'PY'
import numpy as np, sys
sys.path.insert(0, '.')
from unicore_b200 import prostt5_spec as spec
aa, off = spec.synthetic_proteome("config4", n=20000)
with open("/tmp/c4.fa", "w") as f:
    for i in range(len(off) - 1):
        f.write(f">s{i}\n{aa[int(off[i]):int(off[i+1])].tobytes().decode()}\n")
PY
mkdir -p /tmp/c4in && mv /tmp/c4.fa /tmp/c4in/Synth.fa
timeout 600 unicore_b200/bin/unicore-b200 createdb /tmp/c4in /tmp/c4out/db /tmp/p5_full_seed1 --stats-json gpurun_out/cli_config4_20k_8gpu.json 2>&1 | tail -3
cat gpurun_out/cli_config4_20k_8gpu.json
