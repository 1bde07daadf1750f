#!/bin/bash
# ncu launch list of the bench command (contract: per-launch durations; cold-cache and serialised, so shares, not absolutes)
# + a full capture of the bounded-scan NJ kernel at 10,000 nodes' shape of work (6,000 here to keep it short)
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_short.json 2> gpurun_out/r02_bench_short.err || { tail -5 gpurun_out/r02_bench_short.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches_bench_c4.csv python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline --no-e2e > gpurun_out/r02_ncu_launches.log 2>&1; echo launches rc=$?
python tools/nj_once.py 4000 2000000 1 > gpurun_out/r02_nj_once.log 2>&1 || { tail -5 gpurun_out/r02_nj_once.log; exit 1; }
ncu --set full --import-source on --clock-control none -k regex:nj_lazy_kernel -c 1 -f -o gpurun_out/r02_nj_lazy python tools/nj_once.py 4000 2000000 1 > gpurun_out/r02_ncu_nj.log 2>&1; echo ncu rc=$?; tail -2 gpurun_out/r02_ncu_nj.log
