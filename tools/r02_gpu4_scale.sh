#!/bin/bash
# the driver's scaling launch at N = 4 (both arms), as a check of the one N this round had not run
mkdir -p gpurun_out
N=4
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --impl reference --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_ref_n$N.json 2> gpurun_out/r02_bench_ref_n$N.err ) 2>&1 | tail -3; tail -1 gpurun_out/r02_bench_ref_n$N.json | cut -c1-400
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err ) 2>&1 | tail -3; tail -2 gpurun_out/r02_bench_n$N.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_n$N.json') if l.startswith('{')][-1])
print('n$N', d['ms_per_step'], d['stages_ms'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['h2d_gbps_per_rank'], d['config'])
PY
