"""Full-size parity anchors for BASELINE configs C1, C3 and C4, computed ONCE on the CPU by the oracle
(oracle/pf_oracle.c) and committed as small fixtures, so that the `-m gpu` tests can check the GPU's
full-size count matrices and trees without hours of CPU work on the GPU box:

  fullsize_<cfg>.npz  sha256 of the int32 [3, n, n] count matrices (V, D1, D2; C order), and for both metrics
                      (ibs, p) the oracle's NJ join list + branch lengths computed with row sums recomputed
                      FROM SCRATCH every iteration (the specified form) and with the
                      incremental row sums the kernel uses.

    python tests/golden/make_fullsize_fixtures.py [c1 c3 c4]      (c4: ~10 minutes of counts on 8 cores)

Inputs are the seeded synthetic genotypes of SURVEY.md §8(d) (seed 42), generated chunk by chunk.
"""
import hashlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import cpu as oracle  # noqa: E402

CONFIGS = {"c1": (50, 100_000, 0), "c3": (1000, 1_000_000, 0), "c4": (3000, 10_000_000, 0)}
SEED = 42


def counts_of(n, m, miss):
    cache = os.path.join(ROOT, "build", "sim", f"counts_{n}_{m}_{miss}.npy")
    if os.path.exists(cache):
        return np.load(cache).astype(np.int64)
    chunk = 1 << 18
    counts = np.zeros((3, n, n), dtype=np.int64)
    for s0 in range(0, m, chunk):
        k = min(chunk, m - s0)
        codes = oracle.synth_codes(n, s0, k, seed=SEED, miss_ppm=miss)
        lo, hi = oracle.planes64(np.ascontiguousarray(codes.T))
        counts += oracle.pair_counts_bits(lo, hi)
    return counts


def main(names):
    for name in names:
        n, m, miss = CONFIGS[name]
        t0 = time.time()
        counts = counts_of(n, m, miss)
        out = {"n": n, "m": m, "miss_ppm": miss, "seed": SEED,
               "counts_sha256": hashlib.sha256(np.ascontiguousarray(counts.astype(np.int32)).tobytes()).hexdigest()}
        for metric in ("ibs", "p"):
            dist, _ = oracle.normalise(counts, metric)
            out[f"dist_sha256_{metric}"] = hashlib.sha256(np.ascontiguousarray(dist).tobytes()).hexdigest()
            m_inc, l_inc = oracle.nj(dist)
            m_fs, l_fs = oracle.nj(dist, from_scratch=True)
            out[f"merge_incremental_{metric}"], out[f"length_incremental_{metric}"] = m_inc, l_inc
            out[f"merge_from_scratch_{metric}"], out[f"length_from_scratch_{metric}"] = m_fs, l_fs
        np.savez_compressed(os.path.join(HERE, f"fullsize_{name}.npz"), **out)
        print(name, "done in %.1f s" % (time.time() - t0), out["counts_sha256"])


if __name__ == "__main__":
    main(sys.argv[1:] or ["c1", "c3", "c4"])
