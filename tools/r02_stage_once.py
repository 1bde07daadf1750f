"""One pass of the hot path on 3,000 x M sites (default 2,000,000) for ncu: pack -> operand stream -> GEMM -> normalise -> NJ."""
import sys
import torch
sys.path.insert(0, ".")
from phyloforge_b200.engine import Engine
eng = Engine(0)
n, m = 3000, int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
codes = eng.synth_codes(n, 0, m, seed=42)
nw = eng.n_words(m)
planes = eng.alloc_planes(n, nw)
for _ in range(2):
    counts = eng.alloc_counts(n)
    eng.pack_codes(codes, n, planes, nw)
    eng.pair_counts(planes, n, nw, counts)
    dist, _ = eng.normalise(counts, "ibs")
    eng.nj(dist, destroy=True)
torch.cuda.synchronize()
print("ok")
