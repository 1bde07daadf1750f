#!/bin/bash
# final-state evidence: ncu launch list of the default bench command, then the default bench itself
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/final_plain.json 2> gpurun_out/final_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
timeout 400 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/bench_default.json'));print(d['ms_per_step'],d['value'],d['stages_ms'],d['e2e']['ms_per_step'],d['roofline']['frac'],d['clocks'],d['cpu_baseline']['value'])"
