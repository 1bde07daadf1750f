"""Bootstrap support of the NJ tree (the reference asks its tree tool for
`-b 1000`, treetool/script.py:359-360, unseeded). Block bootstrap: the site axis is cut into B contiguous
blocks, a replicate draws B blocks with replacement (pf_bootstrap_weights, seeded), its count matrices
are the weighted sums of the block matrices (pf_bootstrap_counts), and distances + NJ run per replicate
on the GPU. This module holds the host side: which splits of the main tree each replicate tree has."""
from __future__ import annotations

import numpy as np

_MASK = (1 << 64) - 1


def _leaf_keys(n: int):
    """Two independent 64-bit keys per leaf (fixed, SplitMix64 of the leaf index)."""
    def splitmix(x):
        x = (x + 0x9E3779B97F4A7C15) & _MASK
        x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _MASK
        x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _MASK
        return x ^ (x >> 31)
    k1 = np.array([splitmix(2 * i + 1) for i in range(n)], dtype=np.uint64)
    k2 = np.array([splitmix(0x5851F42D4C957F2D ^ (2 * i)) for i in range(n)], dtype=np.uint64)
    return k1, k2


def split_keys(merges: np.ndarray, n: int) -> np.ndarray:
    """merges [R, n-2, 3] (pf_nj join lists) -> [R, n-3] uint64 keys, one per internal edge (the edge
    above node n+t, t < n-3): a commutative 128-bit hash of the leaf set on the side without leaf 0,
    folded to 64 bits. Equal splits give equal keys; different splits collide with probability ~2^-64."""
    merges = np.asarray(merges)
    if merges.ndim == 2:
        merges = merges[None]
    R = merges.shape[0]
    k1, k2 = _leaf_keys(n)
    h1 = np.zeros((R, 2 * n - 3), dtype=np.uint64)
    h2 = np.zeros((R, 2 * n - 3), dtype=np.uint64)
    has0 = np.zeros((R, 2 * n - 3), dtype=bool)
    h1[:, :n] = k1
    h2[:, :n] = k2
    has0[:, 0] = True
    rows = np.arange(R)
    with np.errstate(over="ignore"):
        for t in range(n - 3):
            a = merges[:, t, 0].astype(np.int64)
            b = merges[:, t, 1].astype(np.int64)
            h1[:, n + t] = h1[rows, a] + h1[rows, b]
            h2[:, n + t] = h2[rows, a] + h2[rows, b]
            has0[:, n + t] = has0[rows, a] | has0[rows, b]
        t1 = k1.sum(dtype=np.uint64)
        t2 = k2.sum(dtype=np.uint64)
        a1 = np.where(has0[:, n:], t1 - h1[:, n:], h1[:, n:])
        a2 = np.where(has0[:, n:], t2 - h2[:, n:], h2[:, n:])
        return a1 ^ (a2 * np.uint64(0x9E3779B97F4A7C15))


def split_support(main_merge: np.ndarray, rep_merges: np.ndarray, n: int, threads: int = 0) -> np.ndarray:
    """Percentage of replicate trees that contain each internal edge of the main tree: [n-3] float64,
    entry t belongs to node n+t of the main join list. Computed by the library's host threads
    (csrc/artefacts.cu: pf_bootstrap_support, the same split keys as split_keys below)."""
    if n < 4:
        return np.zeros((0,), dtype=np.float64)
    from . import _lib
    main = np.ascontiguousarray(main_merge, dtype=np.int32)
    reps = np.ascontiguousarray(rep_merges, dtype=np.int32)
    if reps.ndim == 2:
        reps = reps[None]
    if main.shape != (n - 2, 3) or reps.shape[1:] != (n - 2, 3):
        raise ValueError(f"join lists of shape {main.shape} / {reps.shape} for {n} leaves")
    out = np.zeros(n - 3, dtype=np.float64)
    _lib.check(_lib.load().pf_bootstrap_support(main.ctypes.data, reps.ctypes.data, n, reps.shape[0], out.ctypes.data,
                                                threads))
    return out


def split_support_py(main_merge: np.ndarray, rep_merges: np.ndarray, n: int) -> np.ndarray:
    """The same numbers with numpy: the statement the tests hold the library to."""
    if n < 4:
        return np.zeros((0,), dtype=np.float64)
    main = split_keys(main_merge, n)[0]
    reps = split_keys(rep_merges, n)
    # every replicate key looked up in the sorted distinct keys of the main tree at once; a replicate counts once per
    # main edge however often it holds the key (hit is a boolean table)
    uniq, inverse = np.unique(main, return_inverse=True)
    pos = np.minimum(np.searchsorted(uniq, reps), uniq.shape[0] - 1)
    rows, cols = np.nonzero(uniq[pos] == reps)
    hit = np.zeros((reps.shape[0], uniq.shape[0]), dtype=bool)
    hit[rows, pos[rows, cols]] = True
    hits = hit.sum(axis=0, dtype=np.int64)[inverse]
    return 100.0 * hits / max(1, reps.shape[0])
