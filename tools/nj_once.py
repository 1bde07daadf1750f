"""One pf_nj call on a synthetic genotype distance matrix (ncu target).  python tools/nj_once.py n sites [reps]"""
import sys
import torch
sys.path.insert(0, ".")
sys.path.insert(0, "tools")
from phyloforge_b200.engine import Engine
n, m = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
eng = Engine(0)
counts = eng.alloc_counts(n)
chunk = 1 << 20
nw = chunk // 32
planes = eng.alloc_planes(n, nw)
for s0 in range(0, m, chunk):
    k = min(chunk, m - s0)
    codes = eng.synth_codes(n, s0, k, seed=42)
    eng.pack_codes(codes, n, planes, nw, n_sites=k)
    eng.pair_counts(planes, n, nw, counts, word_end=(k + 31) // 32)
dist, _ = eng.normalise(counts, "p")
for _ in range(reps):
    eng.nj(dist)
torch.cuda.synchronize()
print("ok")
