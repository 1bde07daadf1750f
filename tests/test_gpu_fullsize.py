"""BASELINE.json configurations at full size on the GPU, checked through size-independent
properties and exact row-subset comparisons with the oracle (SURVEY.md §8(c) last paragraph)."""
import os
import time

import numpy as np
import pytest
import torch

from oracle import cpu as oracle, treecmp
from phyloforge_b200 import sharding, synth, vcf
from phyloforge_b200.newick import merges_to_newick

pytestmark = pytest.mark.gpu


def _subset_check(engine, n, m, miss_ppm, counts, rows, seed=42):
    """counts[:, rows][:, :, rows] must equal the oracle's counts of those samples over ALL sites."""
    rows = np.asarray(rows, dtype=np.int64)
    sub = oracle.synth_rows(rows, n, 0, m, seed=seed, miss_ppm=miss_ppm)
    lo, hi = oracle.planes64(sub)
    want = oracle.pair_counts_bits(lo, hi)
    idx = torch.from_numpy(rows).to(counts.device)
    got = counts.index_select(1, idx).index_select(2, idx).cpu().numpy().astype(np.int64)
    assert np.array_equal(got, want)


def _counts_in_chunks(engine, n, m, miss_ppm, chunk_sites, seed=42):
    counts = engine.alloc_counts(n)
    nw = chunk_sites // 32
    planes = engine.alloc_planes(n, nw)
    for s0 in range(0, m, chunk_sites):
        k = min(chunk_sites, m - s0)
        codes = engine.synth_codes(n, s0, k, seed=seed, miss_ppm=miss_ppm)
        engine.pack_codes(codes, n, planes, nw, n_sites=k)
        engine.pair_counts(planes, n, nw, counts, word_end=(k + 31) // 32)
        del codes
    return counts


def test_config3_1000x1M_full_matrix_and_tree(engine):
    """configs[2]: 1,000 samples x 1M SNPs, whole count matrices against the oracle, tree RF = 0."""
    n, m = 1000, 1_000_000
    names = [f"s{i}" for i in range(n)]
    codes = engine.synth_codes(n, 0, m, seed=42)
    nw = engine.n_words(m)
    planes = engine.alloc_planes(n, nw)
    engine.pack_codes(codes, n, planes, nw)
    res = engine.tree_from_planes(planes, n, nw, names)
    sm = np.ascontiguousarray(oracle.synth_codes(n, 0, m, seed=42).T)
    want = oracle.pair_counts(sm, method="bits")
    assert np.array_equal(res.counts.cpu().numpy().astype(np.int64), want)
    dist, _ = oracle.normalise(want, "ibs")
    assert np.array_equal(res.dist.cpu().numpy(), dist)
    merge, length = oracle.nj(dist)
    assert np.array_equal(res.merge, merge) and np.array_equal(res.length, length)
    assert treecmp.compare(res.newick, merges_to_newick(merge, length, names)) == (0, 0.0)


def test_config4_3000x10M_properties_and_row_subset(engine):
    """configs[3]: 3,000 x 10M. V == M everywhere (no missing), symmetry, zero diagonal, shard sums ==
    whole (the 1/2/4/8-GPU identity), and a seeded 48-sample subset exact against the oracle."""
    n, m = 3000, 10_000_000
    counts = _counts_in_chunks(engine, n, m, 0, 1 << 21)
    V, D1, D2 = counts[0], counts[1], counts[2]
    assert bool((V == m).all())
    assert torch.equal(D1, D1.T) and torch.equal(D2, D2.T)
    assert bool((torch.diagonal(D1) == 0).all()) and bool((D2 <= D1).all())
    rows = np.sort(np.random.default_rng(4).choice(n, 48, replace=False))
    _subset_check(engine, n, m, 0, counts, rows)
    # sharding identity on a slice of the site axis: 8 word-aligned shards sum to the whole
    m2 = 1_000_003
    whole = _counts_in_chunks(engine, n, m2, 0, 1 << 20)
    parts = engine.alloc_counts(n)
    for r in range(8):
        b, e = sharding.shard_sites(m2, 8, r)
        codes = engine.synth_codes(n, b, e - b, seed=42)
        nw = engine.n_words(e - b)
        planes = engine.alloc_planes(n, nw)
        engine.pack_codes(codes, n, planes, nw)
        engine.pair_counts(planes, n, nw, parts)
    assert torch.equal(whole, parts)
    # tree from the full counts: NJ output is a valid tree over all leaves, bit-identical to the oracle's
    names = [f"s{i}" for i in range(n)]
    res = engine.tree_from_counts(counts, names)
    dist, _ = oracle.normalise(counts.cpu().numpy().astype(np.int64), "ibs")
    merge, length = oracle.nj(dist)
    assert np.array_equal(res.merge, merge) and np.array_equal(res.length, length)


def test_config5_10000x5M_missing_row_subset(engine):
    """configs[4]: 10,000 x 5M with 10% missing: V <= M, symmetric, subset exact, NJ runs at N = 10,000."""
    n, m, miss = 10000, 5_000_000, 100_000
    counts = _counts_in_chunks(engine, n, m, miss, 1 << 20)
    V, D1, D2 = counts[0], counts[1], counts[2]
    assert bool((V <= m).all()) and bool((V > 0).all())
    assert torch.equal(V, V.T) and torch.equal(D1, D1.T) and torch.equal(D2, D2.T)
    assert bool((D2 <= D1).all()) and bool((D1 <= V).all())
    rows = np.sort(np.random.default_rng(5).choice(n, 40, replace=False))
    _subset_check(engine, n, m, miss, counts, rows)
    # the two independent implementations (tcgen05 indicator GEMMs, popcount) agree on whole matrices
    k = 1 << 19
    codes = engine.synth_codes(n, 12345 * 32, k, seed=42, miss_ppm=miss)
    nw = k // 32
    planes = engine.alloc_planes(n, nw)
    engine.pack_codes(codes, n, planes, nw)
    a, b = engine.alloc_counts(n), engine.alloc_counts(n)
    assert engine.pair_counts(planes, n, nw, a, variant=3) == 3 and engine.tc_launches == 2
    assert engine.pair_counts(planes, n, nw, b, variant=2) == 2
    assert torch.equal(a, b)
    del a, b, planes, codes
    names = [f"s{i}" for i in range(n)]
    # NJ at N = 10,000 against the oracle: tests/test_gpu_nj_pinned.py. Here: the general kernel (other scan item
    # width, two grid barriers per iteration, row sums in global memory) agrees with the fast one bit for bit
    from phyloforge_b200._lib import check
    dist, _ = engine.normalise(counts, "ibs")
    fast = engine.nj(dist)
    check(engine.lib.pf_nj_profile(2, None))
    try:
        general = engine.nj(dist)
    finally:
        check(engine.lib.pf_nj_profile(0, None))
    assert torch.equal(fast[0], general[0]) and torch.equal(fast[1], general[1])
    del dist, fast, general
    res = engine.tree_from_counts(counts, names, keep=False)
    tree = treecmp.parse_newick(res.newick)
    assert sorted(treecmp.leaves(tree)) == sorted(names)


def _cfg(tmp_path, section, vcf_path, out, tool):
    p = tmp_path / "run.cfg"
    p.write_text(f"[{section}]\nvcf_file = {vcf_path}\nout_path = {out}\ntree_software = {tool}\nthread = 8\n")
    return str(p)


def test_config1_50x100k_vcf_through_the_entry_point(engine, tmp_path):
    """configs[0]: a real 23 MB VCF through `phyloforge -s`: all_snp.fa equals the reference's
    character table applied to the genotypes; tree equals the oracle's."""
    from phyloforge_b200 import run
    n, m = 50, 100_000
    names = [f"S{i:03d}" for i in range(n)]
    codes = oracle.synth_codes(n, 0, m, seed=42)
    text, ref, alt = synth.snp_vcf_bytes(codes, names)
    path = tmp_path / "c1.vcf"
    path.write_bytes(text)
    out = tmp_path / "out"
    t0 = time.perf_counter()
    run.main(["-s", _cfg(tmp_path, "snp_opt", path, out, "treebest")])
    wall = time.perf_counter() - t0
    want_chars = synth.expected_snp_chars(codes, ref, alt).T
    fa = (out / "all_snp.fa").read_bytes().split(b"\n")
    assert [h[1:].decode() for h in fa[0:-1:2]] == names
    assert b"".join(fa[1::2]) == want_chars.tobytes()
    counts = oracle.pair_counts(np.ascontiguousarray(codes.T), method="bits")
    dist, _ = oracle.normalise(counts, "p")                       # the entry point's default metric
    merge, length = oracle.nj(dist)
    rf, dl = treecmp.compare((out / "SNP_treebest.NHX").read_text(), merges_to_newick(merge, length, names))
    assert rf == 0 and dl <= 1e-9
    print(f"config1 entry point wall {wall:.2f} s (reference converters alone: 2.97 s on 8 cores, BASELINE.md)")


def test_config2_500x100k_sv_vcf_5pct_missing(engine, tmp_path):
    """configs[1]: 500 samples x 100k SV records, 5% missing, through `phyloforge -S` (tree_software = nj)."""
    from phyloforge_b200 import run
    n, m = 500, 100_000
    names = [f"V{i:03d}" for i in range(n)]
    codes = oracle.synth_codes(n, 0, m, seed=7, miss_ppm=50_000)
    text, svtype = synth.sv_vcf_bytes(codes, names)
    path = tmp_path / "c2.vcf"
    path.write_bytes(text)
    out = tmp_path / "out"
    t0 = time.perf_counter()
    run.main(["-S", _cfg(tmp_path, "sv_opt", path, out, "nj")])
    wall = time.perf_counter() - t0
    want = synth.expected_sv_columns(codes, svtype)              # sample-major codes
    lut = np.array([ord("0"), 0, ord("1"), ord(".")], dtype=np.uint8)
    mat = (out / "SV.matrix").read_bytes().split(b"\n")
    assert b"".join(mat[1::2]) == lut[want].tobytes()
    counts = oracle.pair_counts(want, method="bits")
    dist, _ = oracle.normalise(counts, "ibs")
    merge, length = oracle.nj(dist)
    rf, dl = treecmp.compare((out / "SV.treefile").read_text(), merges_to_newick(merge, length, names))
    assert rf == 0 and dl <= 1e-9
    print(f"config2 entry point wall {wall:.2f} s (reference converter alone: 766.5 s, BASELINE.md)")
