#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --workload c4 --variant 3 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/ncu_tc_plain.json 2> gpurun_out/ncu_tc_plain.err && \
ncu --set full --clock-control none --import-source on -k regex:pair_counts_tc -s 1 -c 1 -o gpurun_out/prof_pair_counts_tc -f $CMD > gpurun_out/ncu_tc_full.log 2>&1
echo "full rc=$?"; tail -3 gpurun_out/ncu_tc_full.log
