#!/bin/bash
# 2 GPUs: 1,000 bootstrap replicates through the entry point on worker threads (2 workers, then 8 workers on the 2 devices)
mkdir -p gpurun_out
timeout 600 python tools/r02_vcf_mg.py 2 200000 threads-nccl 2 1000 > gpurun_out/r02_vcf_mg2_threads_boot1000.json 2> gpurun_out/r02_vcf_mg2_threads_boot1000.err; echo "rc=$?"; tail -30 gpurun_out/r02_vcf_mg2_threads_boot1000.json
timeout 600 python tools/r02_vcf_mg.py 8 200000 threads-p2p 2 1000 > gpurun_out/r02_vcf_mg2x4_threads_boot1000.json 2> gpurun_out/r02_vcf_mg2x4_threads_boot1000.err; echo "rc=$?"; tail -22 gpurun_out/r02_vcf_mg2x4_threads_boot1000.json
