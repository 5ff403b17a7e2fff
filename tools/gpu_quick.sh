#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "neighbour or bit_identical or peaked" 2>&1 | tail -8
timeout 600 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "ragged or config2 or batching" 2>&1 | tail -8
