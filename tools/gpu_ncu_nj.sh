#!/bin/bash
# ncu: one full capture of the NJ kernel inside a short default-workload bench
mkdir -p gpurun_out
CMD="python bench.py --workload ${WL:-c4} --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/ncu_nj_plain.json 2> gpurun_out/ncu_nj_plain.err && \
ncu --set full --clock-control none --import-source on -k regex:nj_fast -s 1 -c 1 -o gpurun_out/prof_nj -f $CMD > gpurun_out/ncu_nj.log 2>&1
echo "full rc=$?"; tail -3 gpurun_out/ncu_nj.log | cut -c1-200
