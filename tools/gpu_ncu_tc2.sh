#!/bin/bash
# one ncu --set full capture of the CTA-pair tensor kernel (after the same command has run clean without ncu)
mkdir -p gpurun_out
CMD="python tools/ab_tc.py 3000 4000000 ${1:-2:1:0} noparity"
timeout 200 $CMD > gpurun_out/ncu_tc2_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_tc2_plain.log; exit 1; }
tail -1 gpurun_out/ncu_tc2_plain.log
timeout 500 ncu --set full --clock-control none --import-source on -k regex:pair_counts_tc2 -s 2 -c 1 -o gpurun_out/prof_pair_counts_tc2 -f $CMD > gpurun_out/ncu_tc2_full.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_tc2_full.log
