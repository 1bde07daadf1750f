#!/bin/bash
# round 2: worker threads behind gpus = N on a 2-GPU box: the multi-worker tests (in-process NCCL for threads, process
# group for processes), the converters' throughput on one GPU, then C3-shaped VCF text through the entry point
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_multigpu.py tests/test_gpu_vcf.py -q -x 2>&1 | tail -25 ) > gpurun_out/r02_mg2_threads_tests.log 2>&1; tail -6 gpurun_out/r02_mg2_threads_tests.log
timeout 300 python tools/bench_vcf.py > gpurun_out/r02_vcf_ingest_throughput.json 2> gpurun_out/r02_vcf_ingest.err; echo "rc=$?"; grep -A3 "planes_only\|with_chars" gpurun_out/r02_vcf_ingest_throughput.json | grep "text_GBps"
PHYLOFORGE_B200_TRACE=1 timeout 600 python tools/r02_vcf_mg.py 2 200000 threads-nccl,threads-p2p 2 > gpurun_out/r02_vcf_mg2_threads_200k.json 2> gpurun_out/r02_vcf_mg2_threads_200k.err; echo "rc=$?"; grep "trace" gpurun_out/r02_vcf_mg2_threads_200k.err | tail -34
PHYLOFORGE_B200_TRACE=1 timeout 600 python tools/r02_vcf_mg.py 2 1000000 threads-nccl,threads-p2p 2 > gpurun_out/r02_vcf_mg2_threads_1M.json 2> gpurun_out/r02_vcf_mg2_threads_1M.err; echo "rc=$?"; cat gpurun_out/r02_vcf_mg2_threads_1M.json; grep "trace" gpurun_out/r02_vcf_mg2_threads_1M.err | tail -34
