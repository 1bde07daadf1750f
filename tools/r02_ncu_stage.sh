#!/bin/bash
mkdir -p gpurun_out
CMD="python tools/r02_stage_once.py 2000000"
$CMD > gpurun_out/r02_stage_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'pack_kernel|planes_to_e2m1|pair_counts_tc2|nj_fast' -s 4 -c 4 -o gpurun_out/r02_stage $CMD > gpurun_out/r02_ncu_stage.log 2>&1
echo "ncu rc=$?"; tail -5 gpurun_out/r02_ncu_stage.log; ls -la gpurun_out/*.ncu-rep
