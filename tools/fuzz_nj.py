"""Randomised parity run of the three NJ kernels (bounded scan forced on at every size, fast full scan, general) and of
batches against the CPU oracle: join list and branch lengths bit for bit. Matrix families: uniform, rational with a
common denominator (exact ties), star-plus-noise at several noise levels (near-ties everywhere: the hard case for the
bounds), clustered (populations), additive trees, duplicates, sparse zeros, huge / tiny scales.
    python tools/fuzz_nj.py [seed] [seconds]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import cpu as oracle
from phyloforge_b200._lib import check
from phyloforge_b200.engine import Engine

eng = Engine(0)
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 2024)
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 120.0
t_end = time.time() + budget


def sym(a):
    d = np.triu(a, 1)
    return d + d.T


def family(kind, n):
    if kind == 0:
        return sym(rng.random((n, n)))
    if kind == 1:
        den = int(rng.choice([2, 4, 16, 1024, 1 << 20]))
        return sym(rng.integers(0, den, size=(n, n)).astype(np.float64) / den)
    if kind == 2:                                         # star + noise
        eps = float(rng.choice([0.0, 1e-12, 1e-9, 1e-6, 1e-3, 1e-1]))
        return sym(1.0 + eps * rng.standard_normal((n, n)))
    if kind == 3:                                         # populations
        k = int(rng.integers(2, max(3, n // 5)))
        lab = rng.integers(0, k, size=n)
        base = np.where(lab[:, None] == lab[None, :], 0.3, 0.5)
        return sym(base + float(rng.choice([1e-8, 1e-5, 1e-3])) * rng.random((n, n)))
    if kind == 4:                                         # additive (caterpillar-like), symmetrised
        pos = np.cumsum(rng.random(n))
        h = rng.random(n)
        d = np.abs(pos[:, None] - pos[None, :]) + h[:, None] + h[None, :]
        return sym((d + d.T) / 2)
    if kind == 5:                                         # duplicates
        d = sym(rng.random((n, n)))
        for _ in range(max(1, n // 10)):
            i, j = rng.integers(0, n, size=2)
            if i != j:
                d[:, j] = d[:, i]
                d[j, :] = d[i, :]
                d[i, j] = d[j, i] = 0.0
                d[j, j] = 0.0
        return sym(d)
    if kind == 6:                                         # many zeros
        return sym(rng.random((n, n)) * (rng.random((n, n)) < 0.3))
    scale = float(rng.choice([1e-150, 1e-30, 1e30, 1e150]))
    return sym(rng.random((n, n))) * scale


n_cases = n_batches = 0
joined_total = handed_over = 0
sizes = [3, 4, 5, 8, 9, 15, 16, 17, 31, 33, 63, 64, 65, 100, 127, 129, 200, 255, 257, 300, 400, 513, 700]
import ctypes as C
while time.time() < t_end:
    kind = int(rng.integers(0, 8))
    n = int(rng.choice(sizes))
    dist = family(kind, n)
    want_m, want_l = oracle.nj(dist.copy())
    for mode in (8, 4, 2):
        check(eng.lib.pf_nj_profile(mode, None))
        merge, length = eng.nj(torch.from_numpy(dist).to(eng.device))
        torch.cuda.synchronize()
        out = (C.c_ulonglong * 8)()
        check(eng.lib.pf_nj_profile(0, out))
        if mode == 8 and n >= 8:
            joined_total += int(out[7])
            handed_over += 1 - int(out[7])
        ok = np.array_equal(merge.cpu().numpy(), want_m) and np.array_equal(length.cpu().numpy(), want_l, equal_nan=True)
        assert ok, ("nj", kind, n, mode)
    n_cases += 1
    if n_cases % 7 == 0:
        k = int(rng.integers(2, 20))
        nb = int(rng.choice([8, 20, 64, 130, 257]))
        mats = [family(int(rng.integers(0, 8)), nb) for _ in range(k)]
        check(eng.lib.pf_nj_profile(8, None))
        merge, length = eng.nj_batch(torch.from_numpy(np.stack(mats)).to(eng.device))
        torch.cuda.synchronize()
        check(eng.lib.pf_nj_profile(0, None))
        for j in range(k):
            wm, wl = oracle.nj(mats[j].copy())
            assert np.array_equal(merge[j].cpu().numpy(), wm) and np.array_equal(length[j].cpu().numpy(), wl, equal_nan=True), ("batch", k, nb, j)
        n_batches += 1
print(f"fuzz_nj ok: {n_cases} matrices x 3 kernels (bounded scan joined {joined_total}, handed {handed_over} over), "
      f"{n_batches} batches through the bounded scan")
