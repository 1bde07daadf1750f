#!/bin/bash
# ncu: launch list of a short default-workload bench + one full capture of the pair-count kernel
mkdir -p gpurun_out
CMD="python bench.py --workload ${WL:-c4} --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/ncu_plain2.json 2> gpurun_out/ncu_plain2.err && \
ncu --set full --clock-control none --import-source on -k regex:pair_counts -s 1 -c 1 -o gpurun_out/prof_pair_counts -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
tail -3 gpurun_out/ncu_full.log
