#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "tensor_core or pair_counts_bit_exact or outside" 2>&1 | tail -4
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -3
VARIANTS="3" bash tools/gpu_quick.sh
timeout 600 python bench.py --workload c5 --no-e2e --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "c5 rc=$?"; tail -2 gpurun_out/bench_c5.err
python -c "
import json;d=json.load(open('gpurun_out/bench_c5.json'));print(d['ms_per_step'],d['stages_ms'],d['roofline'],d['config'])"
