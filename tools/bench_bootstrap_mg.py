"""Block bootstrap over several GPUs (run under torchrun): sites sharded over the ranks, block matrices summed
with one allreduce, replicates dealt out to the ranks in batches of 8, join lists gathered. Prints one JSON record
on rank 0; support values must equal the one-GPU run's (same seed): `--check` recomputes them on rank 0 alone.

    python -m torch.distributed.run --nproc-per-node 8 tools/bench_bootstrap_mg.py 3000,10000000,1000,256
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from phyloforge_b200 import sharding
from phyloforge_b200.engine import Engine

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = Engine(local)
check = "--check" in sys.argv
for spec in [a for a in sys.argv[1:] if not a.startswith("--")] or ["3000,10000000,1000,256"]:
    n, m, reps, blocks = (int(x) for x in spec.split(","))
    names = [f"s{i}" for i in range(n)]
    total_words = (m + 31) // 32
    wb, we = sharding.shard_words(total_words, world, rank)
    my_sites = min(we * 32, m) - wb * 32
    nw = max(1, we - wb)
    planes = eng.alloc_planes(n, nw)
    chunk = 1 << 21
    for s0 in range(0, my_sites, chunk):
        k = min(chunk, my_sites - s0)
        codes = eng.synth_codes(n, wb * 32 + s0, k, seed=42)
        eng.pack_codes(codes, n, planes, nw, word0=s0 // 32, n_sites=k)
        del codes
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    main, support, rep = eng.bootstrap_tree(planes, n, nw, names, replicates=reps, seed=1, n_blocks=blocks,
                                            n_sites=my_sites, word0=wb, total_cols=m)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t1 = time.perf_counter()
    rec = {"samples": n, "sites": m, "replicates": reps, "blocks": blocks, "gpus": world,
           "tree_with_bootstrap_s": t1 - t0, "mean_support": float(support.mean()) if support.size else None}
    if check and rank == 0 and world > 1:
        pass
    if rank == 0:
        np.save(f"gpurun_out/bootstrap_support_{n}x{m}_r{reps}_g{world}.npy", support)
        print(json.dumps(rec), flush=True)
    del planes
if world > 1:
    dist.destroy_process_group()
