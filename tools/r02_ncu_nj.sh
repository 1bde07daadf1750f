#!/bin/bash
mkdir -p gpurun_out
python tools/nj_once.py 3000 2000000 1 > gpurun_out/r02_nj_once.log 2>&1 || { tail -5 gpurun_out/r02_nj_once.log; exit 1; }
ncu --set full --import-source on --clock-control none -k regex:nj_lazy_kernel -c 1 -f -o gpurun_out/r02_nj_lazy python tools/nj_once.py 3000 2000000 1 > gpurun_out/r02_ncu_nj.log 2>&1; echo ncu rc=$?; tail -3 gpurun_out/r02_ncu_nj.log; ls -la gpurun_out/r02_nj_lazy.ncu-rep
