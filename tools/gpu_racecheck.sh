#!/bin/bash
# racecheck of the attention kernel with and without the TMA-fetched bias table (full hazard records)
mkdir -p gpurun_out
P5_ATTN_FEAT=14 timeout 900 compute-sanitizer --tool racecheck --print-limit 3 python tools/sanitize_target.py > gpurun_out/racecheck_mask14.txt 2>&1
tail -3 gpurun_out/racecheck_mask14.txt
timeout 900 compute-sanitizer --tool racecheck --print-limit 3 python tools/sanitize_target.py > gpurun_out/racecheck_mask15.txt 2>&1
grep -v "^=========     at\|^=========     by" gpurun_out/racecheck_mask15.txt | head -40
