"""Randomised parity run (not part of the test suite): many shapes of every GPU stage against the oracle."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import cpu as oracle
from phyloforge_b200._lib import check
from phyloforge_b200.engine import Engine

eng = Engine(0)
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 12345)
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 240.0
t_end = time.time() + budget
n_counts = n_nj = n_batch = 0
while time.time() < t_end:
    # ---- pair counts, every variant, random missingness structure
    n = int(rng.choice([2, 3, 17, 63, 64, 65, 127, 128, 129, 191, 192, 193, 255, 256, 257, 300, 383, 384, 385, 450, 640, 777]))
    m = int(rng.integers(1, 9000))
    kind = int(rng.integers(0, 5))
    codes = rng.integers(0, 3, size=(m, n), dtype=np.uint8)
    if kind == 1:
        codes[rng.random(m) < 0.05] = 3                       # sites missing everywhere (still uniform)
    elif kind == 2:
        codes[rng.random((m, n)) < rng.choice([0.001, 0.05, 0.5, 0.95])] = 3
    elif kind == 3:
        codes[:, int(rng.integers(0, n))] = 3                 # one sample all missing
    elif kind == 4:
        codes[int(rng.integers(0, m)), int(rng.integers(0, n))] = 3   # a single missing genotype
    d = torch.from_numpy(codes).to(eng.device)
    nw = eng.n_words(m)
    planes = eng.alloc_planes(n, nw)
    eng.pack_codes(d, n, planes, nw)
    want = oracle.pair_counts(np.ascontiguousarray(codes.T), method="bits")
    cut = int(rng.integers(0, nw + 1))
    for variant in (1, 2, 3):
        c = eng.alloc_counts(n)
        eng.pair_counts(planes, n, nw, c, word_begin=0, word_end=cut, variant=variant)
        eng.pair_counts(planes, n, nw, c, word_begin=cut, word_end=nw, variant=variant)
        got = c.cpu().numpy().astype(np.int64)
        assert np.array_equal(got, want), ("counts", n, m, kind, variant, cut)
    n_counts += 1
    # ---- NJ on rational distances with ties, all kernels
    n = int(rng.integers(3, 700))
    den = int(rng.choice([4, 16, 1024, 1 << 20]))
    a = rng.integers(0, den, size=(n, n)).astype(np.float64) / den
    dist = np.triu(a, 1)
    dist = dist + dist.T
    want_m, want_l = oracle.nj(dist)
    for mode in (0, 2):
        check(eng.lib.pf_nj_profile(mode, None))
        merge, length = eng.nj(torch.from_numpy(dist).to(eng.device))
        check(eng.lib.pf_nj_profile(0, None))
        assert np.array_equal(merge.cpu().numpy(), want_m) and np.array_equal(length.cpu().numpy(), want_l), ("nj", n, den, mode)
    n_nj += 1
    # ---- batched NJ
    k = int(rng.integers(1, 12))
    n = int(rng.integers(3, 200))
    mats = []
    for _ in range(k):
        a = rng.random((n, n))
        dd = (a + a.T) / 2
        np.fill_diagonal(dd, 0)
        mats.append(dd)
    merge, length = eng.nj_batch(torch.from_numpy(np.stack(mats)).to(eng.device))
    for j in range(k):
        wm, wl = oracle.nj(mats[j])
        assert np.array_equal(merge[j].cpu().numpy(), wm) and np.array_equal(length[j].cpu().numpy(), wl), ("batch", k, n, j)
    n_batch += 1
print(f"fuzz ok: {n_counts} count cases x 3 variants, {n_nj} NJ cases x 2 kernels, {n_batch} NJ batches")
