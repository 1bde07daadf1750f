"""Synthetic VCF text for benchmarks and full-size tests: genotype codes
from the counter-based generator are rendered as real VCF records with `GT`-only FORMAT fields.
Host-side, numpy-vectorised; not on any timed path."""
from __future__ import annotations

import numpy as np

_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
_GT = {0: b"0/0", 1: b"0/1", 2: b"1/1", 3: b"./."}


def _header(names) -> bytes:
    return (b"##fileformat=VCFv4.2\n##source=phyloforge_b200.synth\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t"
            + "\t".join(names).encode() + b"\n")


def _gt_block(codes_sitemajor: np.ndarray) -> np.ndarray:
    """[m, n] codes -> [m, n, 4] bytes 'a/b\\t' (the last tab of a row becomes '\\n')."""
    lut = np.zeros((4, 4), dtype=np.uint8)
    for c, s in _GT.items():
        lut[c, :3] = np.frombuffer(s, dtype=np.uint8)
        lut[c, 3] = ord("\t")
    out = lut[codes_sitemajor]
    out[:, -1, 3] = ord("\n")
    return out


def snp_vcf_bytes(codes_sitemajor: np.ndarray, names, seed: int = 42):
    """Returns (vcf bytes, ref[m] u8, alt[m] u8). REF/ALT are distinct bases per site."""
    m, n = codes_sitemajor.shape
    rng = np.random.default_rng(seed)
    ref_i = rng.integers(0, 4, m)
    alt_i = (ref_i + rng.integers(1, 4, m)) % 4
    ref, alt = _BASES[ref_i], _BASES[alt_i]
    pos = np.char.zfill(np.arange(1, m + 1).astype("S9"), 9)
    fixed = np.empty((m, 0), dtype=np.uint8)
    pre = np.frombuffer(b"chr1\t", dtype=np.uint8)
    cols = [np.broadcast_to(pre, (m, pre.size)), pos.view(np.uint8).reshape(m, 9),
            np.broadcast_to(np.frombuffer(b"\t.\t", dtype=np.uint8), (m, 3)), ref[:, None],
            np.full((m, 1), ord("\t"), np.uint8), alt[:, None],
            np.broadcast_to(np.frombuffer(b"\t.\tPASS\t.\tGT\t", dtype=np.uint8), (m, 13))]
    fixed = np.concatenate(cols, axis=1)
    body = np.concatenate([fixed, _gt_block(codes_sitemajor).reshape(m, n * 4)], axis=1)
    return _header(names) + body.tobytes(), ref, alt


def sv_vcf_bytes(codes_sitemajor: np.ndarray, names, seed: int = 42, other_type_every: int = 97):
    """SV records: SVTYPE alternates DEL/INS, every `other_type_every`-th record is DUP (filtered).
    Returns (vcf bytes, svtype[m] in {0 DEL, 1 INS, 2 other})."""
    m, n = codes_sitemajor.shape
    idx = np.arange(m)
    svtype = (idx % 2).astype(np.int64)
    svtype[idx % other_type_every == other_type_every - 1] = 2
    info_lut = np.array([b"SVTYPE=DEL;END=1", b"SVTYPE=INS;END=1", b"SVTYPE=DUP;END=1"], dtype="S16")
    info = info_lut[svtype].view(np.uint8).reshape(m, 16)
    pos = np.char.zfill(np.arange(1, m + 1).astype("S9"), 9)
    cols = [np.broadcast_to(np.frombuffer(b"chr1\t", dtype=np.uint8), (m, 5)), pos.view(np.uint8).reshape(m, 9),
            np.broadcast_to(np.frombuffer(b"\tsv\tN\t<SV>\t.\tPASS\t", dtype=np.uint8), (m, 18)), info,
            np.broadcast_to(np.frombuffer(b"\tGT\t", dtype=np.uint8), (m, 4))]
    fixed = np.concatenate(cols, axis=1)
    body = np.concatenate([fixed, _gt_block(codes_sitemajor).reshape(m, n * 4)], axis=1)
    return _header(names) + body.tobytes(), svtype


def expected_snp_chars(codes_sitemajor: np.ndarray, ref: np.ndarray, alt: np.ndarray) -> np.ndarray:
    """The reference's character for each (site, sample) of a biallelic A/C/G/T site."""
    het = np.zeros((256, 256), dtype=np.uint8)
    for a, b, c in (("A", "C", "M"), ("A", "G", "R"), ("A", "T", "W"), ("C", "G", "S"), ("C", "T", "Y"), ("G", "T", "K")):
        het[ord(a), ord(b)] = het[ord(b), ord(a)] = ord(c)
    out = np.empty(codes_sitemajor.shape, dtype=np.uint8)
    out[:] = ord("-")
    r = np.broadcast_to(ref[:, None], out.shape)
    a = np.broadcast_to(alt[:, None], out.shape)
    h = np.broadcast_to(het[ref, alt][:, None], out.shape)
    out[codes_sitemajor == 0] = r[codes_sitemajor == 0]
    out[codes_sitemajor == 2] = a[codes_sitemajor == 2]
    out[codes_sitemajor == 1] = h[codes_sitemajor == 1]
    return out


def expected_sv_columns(codes_sitemajor: np.ndarray, svtype: np.ndarray) -> np.ndarray:
    """The reference's SV.matrix as 2-bit codes, sample-major [n, columns] (svtree.py:58-120):
    DEL maps allele 0 -> '1', INS maps allele 0 -> '0'; one column when every sample's two
    allele strings are equal (0/0, 1/1, ./.), two otherwise."""
    keep = svtype < 2
    c = codes_sitemajor[keep]
    is_del = (svtype[keep] == 0)[:, None]
    # allele characters as codes: '0' -> 0, '1' -> 2, '.' -> 3
    a1 = np.where(c == 3, 3, np.where(c == 2, 1, 0))           # first allele index (0/1) or missing
    a2 = np.where(c == 3, 3, np.where(c >= 1, 1, 0))           # second allele index; het is "0/1"
    def col(a):
        present = np.where(is_del, 1 - a, a)                   # DEL: allele 0 -> '1'
        return np.where(a == 3, 3, np.where(present == 1, 2, 0)).astype(np.uint8)
    c1, c2 = col(a1), col(a2)
    het_site = (c == 1).any(axis=1)
    ncols = np.where(het_site, 2, 1)
    starts = np.concatenate([[0], np.cumsum(ncols)])
    out = np.empty((codes_sitemajor.shape[1], int(starts[-1])), dtype=np.uint8)
    out[:, starts[:-1]] = c1.T
    out[:, starts[:-1][het_site] + 1] = c2[het_site].T
    return out
