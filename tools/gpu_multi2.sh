#!/bin/bash
# 2-GPU pass: bench.py under torchrun (weak scaling, NCCL all-gather in e2e) and the one-process-per-GPU createdb.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
python -c "from unicore_b200 import synth, prostt5_spec as s; synth.model_dir('/tmp/p5_full_seed1', s.FULL, seed=1)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 6 --warmup 3 2> gpurun_out/bench_n2.err | tee gpurun_out/bench_n2_pass2.json | cut -c1-400
tail -3 gpurun_out/bench_n2.err
python - <<'PY'
import sys
sys.path.insert(0, '.')
from unicore_b200 import prostt5_spec as spec
aa, off = spec.synthetic_proteome("config4", n=4000)
with open("/tmp/c4.fasta", "w") as f:
    for i in range(len(off) - 1):
        f.write(f">unicore_{i:010x}\n{aa[int(off[i]):int(off[i+1])].tobytes().decode()}\n")
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
    -m unicore_b200.createdb_dist /tmp/c4.fasta /tmp/c4db_2 --prostt5-model /tmp/p5_full_seed1 --threads 8 --gpu 1 2>&1 | tail -3 | tee gpurun_out/createdb_dist_n2.txt
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m unicore_b200.createdb_dist /tmp/c4.fasta /tmp/c4db_1 --prostt5-model /tmp/p5_full_seed1 2>&1 | tail -2 | tee -a gpurun_out/createdb_dist_n2.txt
for f in "" _ss _h .index _ss.index _h.index .lookup; do cmp /tmp/c4db_1$f /tmp/c4db_2$f && echo "same $f"; done 2>&1 | tee -a gpurun_out/createdb_dist_n2.txt
