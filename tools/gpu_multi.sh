#!/bin/bash
# multi-GPU bench: N ranks, one per GPU, NCCL allreduce of the count matrices
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench N=$N rc=$?"; tail -5 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json
