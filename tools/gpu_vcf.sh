#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_vcf.py tests/test_gpu_fullsize.py -m gpu -x -q -k "vcf or config1 or config2 or snp or sv" > gpurun_out/pytest_vcf.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_vcf.log
timeout 500 python tools/bench_vcf.py > gpurun_out/bench_vcf.json 2> gpurun_out/bench_vcf.err; echo rc=$?; python -c "
import json
for r in json.load(open('gpurun_out/bench_vcf.json')): print(r['kind'],r['samples'],r['records'],r['text_bytes'],{k:round(v['text_GBps'],2) for k,v in r.items() if isinstance(v,dict)})"; tail -3 gpurun_out/bench_vcf.err
