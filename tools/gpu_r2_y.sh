#!/bin/bash
# fifth attention kernel alone on its SM (one CTA per SM): the softmax time of a 128-key tile without MUFU contention
mkdir -p gpurun_out
P5_ATTN_CTAS=1 timeout 600 python tools/ab_attention.py --iters 10 2>&1 | grep "impl 31\|impl  5" | tee gpurun_out/ab_attention_r2y.txt
