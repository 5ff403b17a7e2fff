#!/bin/bash
# ncu --set full of attention kernel 4 on the long shape (steady state) and on config 2
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_tc4 -s 2 -c 1 -f -o gpurun_out/prof_attn_k4_long \
      python tools/attn_target.py 4 long > gpurun_out/prof_attn_k4_long.log 2>&1
tail -2 gpurun_out/prof_attn_k4_long.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_tc4 -s 2 -c 1 -f -o gpurun_out/prof_attn_k4_c2 \
      python tools/attn_target.py 4 config2 > gpurun_out/prof_attn_k4_c2.log 2>&1
tail -2 gpurun_out/prof_attn_k4_c2.log
