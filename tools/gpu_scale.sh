#!/bin/bash
# scaling runs on one box: N = 8, 4, 2 (1 rank per GPU, NCCL); SCALINGS="strong weak"
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for SC in ${SCALINGS:-strong weak}; do
for N in ${NS:-8 4 2}; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) \
  bench.py --gpus $N --steps 5 --warmup 3 --scaling $SC > gpurun_out/bench_${SC}_n$N.json 2> gpurun_out/bench_${SC}_n$N.err
echo "bench $SC N=$N rc=$?"; tail -2 gpurun_out/bench_${SC}_n$N.err | cut -c1-300; python -c "
import json;d=json.load(open('gpurun_out/bench_${SC}_n$N.json'));print('$SC N=$N',d['ms_per_step'],d['value'],d['stages_ms'],'e2e',d['e2e']['ms_per_step'],d['e2e']['value'])"
done
done
