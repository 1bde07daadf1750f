#!/bin/bash
# round 2, 8 GPUs: gpus = N on worker threads - the multi-worker tests over in-process NCCL on 8 devices, then C3 VCF text
# (1,000 x 1M, 4 GB) through `phyloforge -s` with gpus = 1 and gpus = 8 (threads, NCCL / peer-copy reduce), without and
# with 1,000 bootstrap replicates (200k-record text for those: the replicates, not the conversion, are the work)
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_multigpu.py -q -x 2>&1 | tail -15 ) > gpurun_out/r02_mg8_threads_tests.log 2>&1; tail -5 gpurun_out/r02_mg8_threads_tests.log
PHYLOFORGE_B200_TRACE=1 timeout 600 python tools/r02_vcf_mg.py 8 1000000 threads-nccl,threads-p2p 3 > gpurun_out/r02_vcf_mg8_threads.json 2> gpurun_out/r02_vcf_mg8_threads.err; echo "rc=$?"; tail -48 gpurun_out/r02_vcf_mg8_threads.json
timeout 600 python tools/r02_vcf_mg.py 8 200000 threads-nccl 2 1000 > gpurun_out/r02_vcf_mg8_threads_boot1000.json 2> gpurun_out/r02_vcf_mg8_threads_boot1000.err; echo "rc=$?"; tail -30 gpurun_out/r02_vcf_mg8_threads_boot1000.json
