#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "nj or end_to_end or smoke" 2>&1 | tail -6
timeout 300 python tools/nj_profile.py 2>&1 | tail -12
