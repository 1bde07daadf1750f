#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -s --durations=10 > gpurun_out/pytest_fullsize.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/pytest_fullsize.log
