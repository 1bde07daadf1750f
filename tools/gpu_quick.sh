#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for v in 1 2; do
timeout 600 python bench.py --variant $v --no-e2e --no-cpu-baseline --steps 3 > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err; echo "bench v$v rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_v$v.json'));print(d['ms_per_step'],d['stages_ms'],d['roofline']['frac'],d['clocks'])"
done
