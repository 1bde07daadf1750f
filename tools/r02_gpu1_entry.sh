#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_vcf.py tests/test_gpu_multigpu.py -q -x 2>&1 | tail -15 ) > gpurun_out/r02_vcf_tests.log 2>&1; tail -5 gpurun_out/r02_vcf_tests.log
timeout 600 python tools/r02_entry_profile.py 1000000 > gpurun_out/r02_entry_profile_c3.json 2> gpurun_out/r02_entry_profile_c3.err; echo "rc=$?"; cat gpurun_out/r02_entry_profile_c3.json; tail -3 gpurun_out/r02_entry_profile_c3.err
timeout 300 python tools/bench_vcf.py > gpurun_out/r02_vcf_ingest_throughput.json 2> gpurun_out/r02_vcf_ingest.err; echo "rc=$?"; grep "text_GBps" gpurun_out/r02_vcf_ingest_throughput.json
