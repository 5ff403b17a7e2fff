#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "table" 2>&1 | tail -4 | tee gpurun_out/test_table_poison.txt
