#!/bin/bash
# round 2: worker threads behind gpus = N on a 2-GPU box: the multi-worker tests (in-process NCCL for threads, process
# group for processes), then C3-shaped VCF text through the entry point for every worker model
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_multigpu.py -q -x 2>&1 | tail -25 ) > gpurun_out/r02_mg2_threads_tests.log 2>&1; tail -6 gpurun_out/r02_mg2_threads_tests.log
PHYLOFORGE_B200_TRACE=1 timeout 600 python tools/r02_vcf_mg.py 2 200000 threads-nccl,threads-p2p,processes 2 > gpurun_out/r02_vcf_mg2_threads_200k.json 2> gpurun_out/r02_vcf_mg2_threads_200k.err; echo "rc=$?"; cat gpurun_out/r02_vcf_mg2_threads_200k.json; grep -i "trace\|error\|Traceback" gpurun_out/r02_vcf_mg2_threads_200k.err | tail -40
timeout 600 python tools/r02_vcf_mg.py 2 1000000 threads-nccl,threads-p2p 2 > gpurun_out/r02_vcf_mg2_threads_1M.json 2> gpurun_out/r02_vcf_mg2_threads_1M.err; echo "rc=$?"; cat gpurun_out/r02_vcf_mg2_threads_1M.json; tail -5 gpurun_out/r02_vcf_mg2_threads_1M.err
