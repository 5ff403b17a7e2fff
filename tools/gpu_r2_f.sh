#!/bin/bash
# Round 2: full GPU suite on a 2-GPU box after the NCCL load-order fix, then smoke.
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | grep -v "^$" > gpurun_out/gputest_r2f.txt; tail -15 gpurun_out/gputest_r2f.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
