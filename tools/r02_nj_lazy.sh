#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/nj_lazy_check.py "$@" > gpurun_out/r02_nj_lazy.txt 2> gpurun_out/r02_nj_lazy.err; echo rc=$?; cat gpurun_out/r02_nj_lazy.txt; tail -20 gpurun_out/r02_nj_lazy.err
