#!/bin/bash
# one ncu --set full capture of the tensor-core GEMM kernel inside the bench command (after it ran clean without ncu),
# and the launch list of the same command
mkdir -p gpurun_out
CMD="python bench.py --workload c4 --variant 3 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/ncu_tc_plain.json 2> gpurun_out/ncu_tc_plain.err || { echo "plain run failed"; tail -5 gpurun_out/ncu_tc_plain.err; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:pair_counts_tc2_kernel -s 2 -c 1 -o gpurun_out/prof_pair_counts_tc -f $CMD > gpurun_out/ncu_tc_full.log 2>&1
echo "full rc=$?"; tail -3 gpurun_out/ncu_tc_full.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
