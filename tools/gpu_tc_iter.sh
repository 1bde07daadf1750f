#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "tensor_core or pair_counts_bit_exact" 2>&1 | tail -4
VARIANTS="3" bash tools/gpu_quick.sh
