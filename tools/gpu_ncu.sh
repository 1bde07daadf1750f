#!/bin/bash
# ncu: launch list of a short default-workload bench + one full capture of the pair-count GEMM kernel
mkdir -p gpurun_out
CMD="python bench.py --workload ${WL:-c4} --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
# matches per step: the XX/HH launch, then the (early-exit) VV/HV/VH launch: skip one step, take the XX/HH launch
$CMD > gpurun_out/ncu_plain2.json 2> gpurun_out/ncu_plain2.err && \
ncu --set full --clock-control none --import-source on -k regex:pair_counts_tc -s 2 -c 1 -o gpurun_out/prof_pair_counts_tc -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
tail -3 gpurun_out/ncu_full.log
