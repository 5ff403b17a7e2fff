#!/bin/bash
mkdir -p gpurun_out
timeout 120 build/microbench_umma 2>&1 | tee gpurun_out/microbench_umma.txt
