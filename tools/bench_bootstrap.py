"""Time the block bootstrap (Engine.bootstrap_tree) on synthetic genotypes: n m replicates blocks."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from phyloforge_b200.engine import Engine

eng = Engine(0)
out = []
for spec in (sys.argv[1:] or ["1000,1000000,100,256", "3000,10000000,100,256"]):
    n, m, reps, blocks = (int(x) for x in spec.split(","))
    names = [f"s{i}" for i in range(n)]
    nw = eng.n_words(m)
    planes = eng.alloc_planes(n, nw)
    chunk = 1 << 21
    for s0 in range(0, m, chunk):                       # pack in chunks: the byte codes of 3,000 x 10M are 30 GB
        k = min(chunk, m - s0)
        codes = eng.synth_codes(n, s0, k, seed=42)
        eng.pack_codes(codes, n, planes, nw, word0=s0 // 32, n_sites=k)
        del codes
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    plain = eng.tree_from_planes(planes, n, nw, names)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    main, support, _ = eng.bootstrap_tree(planes, n, nw, names, replicates=reps, seed=1, n_blocks=blocks)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    rec = {"samples": n, "sites": m, "replicates": reps, "blocks": blocks, "tree_only_s": t1 - t0,
           "tree_with_bootstrap_s": t2 - t1, "seconds_per_replicate": (t2 - t1 - (t1 - t0)) / reps,
           "mean_support": float(support.mean()), "same_main_tree": bool((plain.merge == main.merge).all())}
    print(json.dumps(rec), flush=True)
    out.append(rec)
    del planes
json.dump(out, open("gpurun_out/bootstrap_timing.json", "w"), indent=1)
