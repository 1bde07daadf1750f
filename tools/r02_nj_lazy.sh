#!/bin/bash
mkdir -p gpurun_out
if [ "$1" = "--tests" ]; then shift; python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "nj" 2>&1 | tail -5; fi
timeout 600 python tools/nj_lazy_check.py "$@" > gpurun_out/r02_nj_lazy.txt 2> gpurun_out/r02_nj_lazy.err; echo rc=$?; cat gpurun_out/r02_nj_lazy.txt; tail -20 gpurun_out/r02_nj_lazy.err
