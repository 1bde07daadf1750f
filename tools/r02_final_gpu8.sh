#!/bin/bash
# round 2, final state on eight GPUs: the bench under torchrun (strong default + weak + C5 extras), 1,000 bootstrap
# replicates of the 3,000 x 10M tree over the ranks, gpus = 8 on worker threads through the entry point (C3 VCF text,
# and 1,000 bootstrap replicates of a 1,000 x 200k tree)
mkdir -p gpurun_out
N=8
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err ) 2>&1 | tail -3; tail -3 gpurun_out/r02_bench_n$N.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_n$N.json') if l.startswith('{')][-1])
print('n$N', d['ms_per_step'], d['stages_ms'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['h2d_gbps_per_rank'])
for k,v in d['extra'].items(): print(k, v['ms_per_step'], v['stages_ms'], (v.get('e2e') or {}).get('ms_per_step'))
PY
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 tools/bench_bootstrap_mg.py 3000,10000000,1000,256 > gpurun_out/r02_bootstrap_mg$N.json 2> gpurun_out/r02_bootstrap_mg$N.err ) 2>&1 | tail -3; tail -2 gpurun_out/r02_bootstrap_mg$N.err; cat gpurun_out/r02_bootstrap_mg$N.json
timeout 600 python tools/r02_vcf_mg.py 8 1000000 threads-nccl 3 > gpurun_out/r02_vcf_mg8_threads.json 2> gpurun_out/r02_vcf_mg8_threads.err; echo "rc=$?"; tail -30 gpurun_out/r02_vcf_mg8_threads.json
timeout 600 python tools/r02_vcf_mg.py 8 200000 threads-nccl 2 1000 > gpurun_out/r02_vcf_mg8_threads_boot1000.json 2> gpurun_out/r02_vcf_mg8_threads_boot1000.err; echo "rc=$?"; tail -30 gpurun_out/r02_vcf_mg8_threads_boot1000.json
