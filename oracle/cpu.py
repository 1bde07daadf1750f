"""TEST INFRASTRUCTURE ONLY — ctypes wrapper of oracle/pf_oracle.c plus tiny numpy
restatements used to cross-check the C code itself. See oracle/__init__.py."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "libpf_oracle.so"
TILE = 64

_lib = None


def build(force: bool = False) -> Path:
    src = HERE / "pf_oracle.c"
    if force or not LIB.exists() or LIB.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "-B", "libpf_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(LIB))
        _lib.pfo_num_threads.restype = C.c_int
        _lib.pfo_normalise.restype = C.c_int64
        _lib.pfo_nj.restype = C.c_int
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def num_threads() -> int:
    return int(lib().pfo_num_threads())


def set_num_threads(n: int) -> int:
    """OpenMP threads of the C oracle (torchrun exports OMP_NUM_THREADS=1 to its workers). Returns the count in use."""
    lib().pfo_set_num_threads(C.c_int(int(n)))
    return num_threads()


# ------------------------------------------------------------------ synthetic genotypes
def synth_codes(n_samples: int, site0: int, n_sites: int, seed: int = 42, n_pops: int | None = None,
                miss_ppm: int = 0) -> np.ndarray:
    """Site-major codes [n_sites, n_samples] u8 (same generator as pf_synth_codes)."""
    if n_pops is None:
        n_pops = default_pops(n_samples)
    out = np.empty((n_sites, n_samples), dtype=np.uint8)
    lib().pfo_synth_codes(_p(out), C.c_int64(n_samples), C.c_int64(n_samples), C.c_int64(site0),
                          C.c_int64(n_sites), C.c_uint64(seed), C.c_int(n_pops), C.c_uint32(miss_ppm))
    return out


def bootstrap_weights(seed: int, replicate0: int, n_reps: int, n_blocks: int) -> np.ndarray:
    """[n_reps, n_blocks] int32 block multiplicities (same draws as pf_bootstrap_weights)."""
    out = np.zeros((n_reps, n_blocks), dtype=np.int32)
    lib().pfo_bootstrap_weights(C.c_uint64(seed), C.c_int64(replicate0), C.c_int(n_reps), C.c_int(n_blocks), _p(out))
    return out


def bootstrap_counts(blocks: np.ndarray, weights: np.ndarray) -> np.ndarray:
    """blocks [B, 3, n, n] (int), weights [R, B] -> replicate count matrices [R, 3, n, n] int64."""
    return np.tensordot(weights.astype(np.int64), blocks.astype(np.int64), axes=(1, 0))


def synth_rows(rows, n_samples: int, site0: int, n_sites: int, seed: int = 42,
               n_pops: int | None = None, miss_ppm: int = 0) -> np.ndarray:
    """Sample-major codes [len(rows), n_sites] for a subset of samples."""
    if n_pops is None:
        n_pops = default_pops(n_samples)
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    out = np.empty((len(rows), n_sites), dtype=np.uint8)
    lib().pfo_synth_rows(_p(out), _p(rows), C.c_int64(len(rows)), C.c_int64(n_sites),
                         C.c_int64(n_samples), C.c_int64(site0), C.c_int64(n_sites),
                         C.c_uint64(seed), C.c_int(n_pops), C.c_uint32(miss_ppm))
    return out


def default_pops(n_samples: int) -> int:
    """K = max(2, N/25) sub-populations."""
    return max(2, n_samples // 25)


# ------------------------------------------------------------------ pack
def pack_canonical(codes_sm: np.ndarray) -> np.ndarray:
    """Sample-major codes [n, m] -> canonical 2-bit matrix [n, stride] (stride % 128 == 0)."""
    codes_sm = np.ascontiguousarray(codes_sm, dtype=np.uint8)
    n, m = codes_sm.shape
    stride = max(128, ((m + 3) // 4 + 127) // 128 * 128)
    out = np.empty((n, stride), dtype=np.uint8)
    lib().pfo_pack_canonical(_p(codes_sm), C.c_int64(n), C.c_int64(m), C.c_int64(m), _p(out),
                             C.c_int64(stride))
    return out


def pack_planes(codes_sm: np.ndarray, n_words: int | None = None) -> np.ndarray:
    """Sample-major codes [n, m] -> the kernels' tiled bit-plane buffer (uint32, flat)."""
    codes_sm = np.ascontiguousarray(codes_sm, dtype=np.uint8)
    n, m = codes_sm.shape
    if n_words is None:
        n_words = (m + 31) // 32
    n_tiles = (n + TILE - 1) // TILE
    out = np.empty(n_tiles * n_words * 2 * TILE, dtype=np.uint32)
    lib().pfo_pack_planes(_p(codes_sm), C.c_int64(n), C.c_int64(m), C.c_int64(m), _p(out),
                          C.c_int64(n_words))
    return out


def shift_planes(planes: np.ndarray, n_samples: int, n_sites: int, shift: int) -> np.ndarray:
    """Restatement of pf_planes_shift on the tiled layout: site s of the source becomes site s + shift; the vacated
    low bits and everything behind the last site are code 3 (all ones). planes: pack_planes(...) of n_sites sites."""
    n_tiles = (n_samples + TILE - 1) // TILE
    nw_src = (n_sites + 31) // 32
    nw_dst = (n_sites + shift + 31) // 32
    src = np.asarray(planes, dtype=np.uint32).reshape(n_tiles, nw_src, 2 * TILE)
    bits = np.unpackbits(src.view(np.uint8).reshape(n_tiles, nw_src, 2 * TILE, 4), axis=-1, bitorder="little")
    bits = bits.reshape(n_tiles, nw_src, 2 * TILE, 32).transpose(0, 2, 1, 3).reshape(n_tiles, 2 * TILE, nw_src * 32)
    out = np.ones((n_tiles, 2 * TILE, nw_dst * 32), dtype=np.uint8)
    out[:, :, shift:shift + n_sites] = bits[:, :, :n_sites]
    out = out.reshape(n_tiles, 2 * TILE, nw_dst, 32).transpose(0, 2, 1, 3)
    words = np.packbits(out, axis=-1, bitorder="little").reshape(n_tiles, nw_dst, 2 * TILE, 4)
    return np.ascontiguousarray(words).view(np.uint32).reshape(-1)


# ------------------------------------------------------------------ counts / distance
def pair_counts(codes_sm: np.ndarray, method: str = "auto") -> np.ndarray:
    """Sample-major codes [n, m] -> int64 [3, n, n] (V, D1, D2)."""
    codes_sm = np.ascontiguousarray(codes_sm, dtype=np.uint8)
    n, m = codes_sm.shape
    out = np.zeros((3, n, n), dtype=np.int64)
    if method == "auto":
        method = "bytes" if n * n * m <= 2_000_000_000 else "bits"
    if method == "bytes":
        lib().pfo_pair_counts_bytes(_p(codes_sm), C.c_int64(n), C.c_int64(m), C.c_int64(m), _p(out),
                                    C.c_int64(n))
    elif method == "bits":
        lo, hi = planes64(codes_sm)
        pair_counts_bits(lo, hi, out)
    else:
        raise ValueError(method)
    return out


def planes64(codes_sm: np.ndarray):
    codes_sm = np.ascontiguousarray(codes_sm, dtype=np.uint8)
    n, m = codes_sm.shape
    nw = (m + 63) // 64
    lo = np.empty((n, nw), dtype=np.uint64)
    hi = np.empty((n, nw), dtype=np.uint64)
    lib().pfo_planes64_from_codes(_p(codes_sm), C.c_int64(n), C.c_int64(m), C.c_int64(m), _p(lo), _p(hi),
                                  C.c_int64(nw))
    return lo, hi


def pair_counts_bits(lo: np.ndarray, hi: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
    n, nw = lo.shape
    if out is None:
        out = np.zeros((3, n, n), dtype=np.int64)
    lib().pfo_pair_counts_bits(_p(lo), _p(hi), C.c_int64(n), C.c_int64(nw), _p(out), C.c_int64(n))
    return out


def pair_counts_numpy(codes_sm: np.ndarray) -> np.ndarray:
    """Independent numpy restatement (tiny inputs) used to check the C oracle itself."""
    c = np.asarray(codes_sm, dtype=np.int64)
    n = c.shape[0]
    out = np.zeros((3, n, n), dtype=np.int64)
    valid = c != 3
    for i in range(n):
        both = valid[i][None, :] & valid
        out[0, i] = both.sum(1)
        out[1, i] = (both & (c[i][None, :] != c)).sum(1)
        out[2, i] = (both & (np.abs(c[i][None, :] - c) == 2)).sum(1)
    return out


_ALLELE_PAIRS = {"A": (0, 0), "C": (1, 1), "G": (2, 2), "T": (3, 3), "M": (0, 1), "R": (0, 2), "W": (0, 3), "S": (1, 2),
                 "Y": (1, 3), "K": (2, 3), "a": (0, 4), "c": (1, 4), "g": (2, 4), "t": (3, 4)}


def char_pair_counts(chars_sm: np.ndarray) -> np.ndarray:
    """int64 [3, n, n] (V, D1, D2) of a sample-major CHARACTER matrix as treetool/vcftree.py:159-171 writes it
    (all_snp.fa, the input of `treebest nj`, treetool/script.py:359-360). Every character stands for one unordered
    allele pair over {A, C, G, T, *} (vcftree.py:26-65); '-', 'N' and '.' are missing. A column is valid for a pair
    when neither character is missing; D1 = valid and the characters differ (mismatch model); D2 = valid and the two
    genotypes share no allele, so that D1 + D2 = 2 - (alleles shared, with multiplicity), the allele-sharing
    distance. On columns that fit the 2-bit code this equals pair_counts of the codes."""
    c = np.asarray(chars_sm, dtype=np.uint8)
    n, m = c.shape
    missing = np.isin(c, np.frombuffer(b"-N.", dtype=np.uint8))
    copies = np.zeros((n, m, 5), dtype=np.int64)
    known = missing.copy()
    for ch, (a1, a2) in _ALLELE_PAIRS.items():
        sel = c == ord(ch)
        copies[sel, a1] += 1
        copies[sel, a2] += 1
        known |= sel
    if not known.all():
        raise ValueError("a character that stands for no genotype: " + repr(bytes(np.unique(c[~known])).decode()))
    valid = ~missing
    out = np.zeros((3, n, n), dtype=np.int64)
    for i in range(n):
        both = valid[i][None, :] & valid
        shared = np.minimum(copies[i][None, :, :], copies).sum(2)
        differ = c[i][None, :] != c
        assert np.array_equal(both & differ, both & (shared < 2))      # one character per allele pair
        out[0, i] = both.sum(1)
        out[1, i] = (both & differ).sum(1)
        out[2, i] = (both & (shared == 0)).sum(1)
    return out


def char_mismatch_counts(chars_sm: np.ndarray) -> np.ndarray:
    """int64 [2, n, n] (V, D1) as above for ANY characters ('-', 'N', '.' missing): the mismatch model alone."""
    c = np.asarray(chars_sm, dtype=np.uint8)
    n = c.shape[0]
    valid = ~np.isin(c, np.frombuffer(b"-N.", dtype=np.uint8))
    out = np.zeros((2, n, n), dtype=np.int64)
    for i in range(n):
        both = valid[i][None, :] & valid
        out[0, i] = both.sum(1)
        out[1, i] = (both & (c[i][None, :] != c)).sum(1)
    return out


def normalise(counts: np.ndarray, metric: str = "ibs"):
    """int64 [3, n, n] -> (fp64 [n, n], number of pairs with V == 0 counted over i != j)."""
    counts = np.ascontiguousarray(counts, dtype=np.int64)
    n = counts.shape[1]
    dist = np.empty((n, n), dtype=np.float64)
    nz = lib().pfo_normalise(_p(counts), C.c_int64(n), C.c_int64(n), C.c_int(metric_id(metric)), _p(dist),
                             C.c_int64(n))
    return dist, int(nz)


def metric_id(metric) -> int:
    if metric in (0, "p", "pdist", "p-distance"):
        return 0
    if metric in (1, "ibs", "IBS"):
        return 1
    raise ValueError(f"unknown metric {metric!r}")


# ------------------------------------------------------------------ NJ
def nj(dist: np.ndarray, from_scratch: bool = False):
    """fp64 [n, n] -> (merge int32 [n-2, 3], len fp64 [n-2, 3]); see pf_nj in the header.
    from_scratch=True recomputes all row sums every iteration (checker of the incremental sums)."""
    d = np.array(dist, dtype=np.float64, order="C", copy=True)
    n = d.shape[0]
    if n < 3:
        raise ValueError("NJ needs n >= 3")
    merge = np.zeros((n - 2, 3), dtype=np.int32)
    length = np.zeros((n - 2, 3), dtype=np.float64)
    fn = lib().pfo_nj_from_scratch if from_scratch else lib().pfo_nj
    rc = fn(_p(d), C.c_int64(n), C.c_int64(n), _p(merge), _p(length))
    if rc != 0:
        raise RuntimeError(f"pfo_nj failed: {rc}")
    return merge, length


def nj_numpy(dist: np.ndarray):
    """Independent restatement of the same NJ spec in plain Python/numpy (tiny n): row sums
    in the 32-lane strided + butterfly order once, then updated incrementally; Q argmin ties ->
    lowest (i, j), merged node in slot i, last slot moved into slot j. Used to check pfo_nj."""
    d = np.array(dist, dtype=np.float64, copy=True)
    n = d.shape[0]
    node = list(range(n))
    merge = np.zeros((n - 2, 3), dtype=np.int32)
    length = np.zeros((n - 2, 3), dtype=np.float64)
    na, t = n, 0

    def rowsum(row):
        p = [np.float64(0.0)] * 32
        for lane in range(32):
            acc = np.float64(0.0)
            for k in range(lane, len(row), 32):
                acc = acc + row[k]
            p[lane] = acc
        off = 16
        while off >= 1:
            p = [p[lane] + p[lane ^ off] for lane in range(32)]
            off >>= 1
        return p[0]

    r = [rowsum(d[a, :na]) for a in range(na)]
    while na > 3:
        n2 = np.float64(na - 2)
        best, ba, bb = np.inf, 0, 1
        for a in range(na):
            for b in range(a + 1, na):
                q = n2 * d[a, b]
                q = q - r[a]
                q = q - r[b]
                if q < best:
                    best, ba, bb = q, a, b
        dab = d[ba, bb]
        ra, rb = r[ba], r[bb]
        tt = (ra - rb) / (np.float64(2.0) * n2)
        la = np.float64(0.5) * dab + tt
        lb = dab - la
        merge[t] = (node[ba], node[bb], 0)
        length[t] = (la, lb, 0.0)
        for k in range(na):
            if k in (ba, bb):
                continue
            dak, dbk = d[ba, k], d[bb, k]
            v = np.float64(0.5) * ((dak + dbk) - dab)
            r[k] = ((r[k] - dak) - dbk) + v
            d[ba, k] = v
            d[k, ba] = v
        r[ba] = np.float64(0.5) * ((ra + rb) - np.float64(na) * dab)
        d[ba, ba] = 0.0
        last = na - 1
        if bb != last:
            d[bb, :na] = d[last, :na].copy()
            d[:na, bb] = d[:na, last].copy()
            d[bb, bb] = 0.0
            node[bb] = node[last]
            r[bb] = r[last]
        node[ba] = n + t
        na -= 1
        t += 1
    d01, d02, d12 = d[0, 1], d[0, 2], d[1, 2]
    merge[t] = (node[0], node[1], node[2])
    length[t] = (0.5 * ((d01 + d02) - d12), 0.5 * ((d01 + d12) - d02), 0.5 * ((d02 + d12) - d01))
    return merge, length
