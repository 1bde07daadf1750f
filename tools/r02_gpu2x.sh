#!/bin/bash
# round 2, two GPUs: gpus = 2 through the entry point over NCCL, the bench under torchrun (strong default + extras),
# the reference arm under torchrun (OMP threads), 2-GPU VCF text ingestion throughput
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multigpu.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r02_mg2_tests.log; tail -5 gpurun_out/r02_mg2_tests.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err ) 2>&1 | tail -3; echo "bench n2 rc=$?"; tail -3 gpurun_out/r02_bench_n2.err
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_ref_n2.json 2> gpurun_out/r02_bench_ref_n2.err ) 2>&1 | tail -3
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n2.json'))
print('n2', d['ms_per_step'], d['stages_ms'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['h2d_gbps_per_rank'])
for k,v in d['extra'].items(): print(k, v['ms_per_step'], v['stages_ms'], (v.get('e2e') or {}).get('ms_per_step'))
r=json.load(open('gpurun_out/r02_bench_ref_n2.json'))
print('ref', r['value'], r['cpu_baseline']['cores'], r['wall_s'], r['config']==d['config'])
PY
python tools/r02_vcf_mg.py 2 > gpurun_out/r02_vcf_mg2.json 2> gpurun_out/r02_vcf_mg2.err; tail -3 gpurun_out/r02_vcf_mg2.err; cat gpurun_out/r02_vcf_mg2.json
