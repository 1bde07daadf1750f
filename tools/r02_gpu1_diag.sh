#!/bin/bash
# one GPU: the multi-worker tests; gpus = 8 worker threads sharing the device, traced per chunk
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_multigpu.py tests/test_gpu_vcf.py -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/r02_mg_tests.log 2>&1; tail -5 gpurun_out/r02_mg_tests.log
PHYLOFORGE_B200_TRACE=1 timeout 600 python tools/r02_vcf_mg.py 8 1000000 threads-p2p 3 > gpurun_out/r02_vcf_diag.json 2> gpurun_out/r02_vcf_diag.err; echo "rc=$?"; tail -34 gpurun_out/r02_vcf_diag.json; grep "trace rank [0]" gpurun_out/r02_vcf_diag.err | tail -18
