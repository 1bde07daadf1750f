#!/bin/bash
mkdir -p gpurun_out
for v in ${VARIANTS:-2 3}; do
timeout 600 python bench.py --variant $v --no-e2e --no-cpu-baseline --steps 3 > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err; echo "bench v$v rc=$?"; tail -2 gpurun_out/bench_v$v.err; python -c "
import json;d=json.load(open('gpurun_out/bench_v$v.json'));print(d['ms_per_step'],d['stages_ms'],d['roofline']['frac'],d['clocks'])"
done
