#!/bin/bash
# ncu of the two tcgen05 attention kernels (isolated launch, config 2 shape)
set -x
mkdir -p gpurun_out
for impl in 2 31; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 2 -c 1 -f -o gpurun_out/prof_attn_impl$impl \
      python tools/attn_target.py $impl config2 > gpurun_out/prof_attn_impl$impl.log 2>&1
  tail -2 gpurun_out/prof_attn_impl$impl.log
done
