"""Parity tests proper: every CUDA stage, called through the C ABI, against the CPU oracle
on the same seeded inputs (bit-exact for packed matrices, integer counts, distances, NJ)."""
import numpy as np
import pytest
import torch

from oracle import cpu as oracle, treecmp
from phyloforge_b200.newick import merges_to_newick
from phyloforge_b200 import _lib

pytestmark = pytest.mark.gpu


def _random_codes(rng, n, m, miss):
    c = rng.integers(0, 3, size=(m, n), dtype=np.uint8)          # site-major
    if miss > 0:
        c[rng.random((m, n)) < miss] = 3
    return c


def _pack(engine, codes_sitemajor, ld=None):
    m, n = codes_sitemajor.shape
    ld = ld or n
    buf = np.full((m, ld), 3, dtype=np.uint8)
    buf[:, :n] = codes_sitemajor
    d = torch.from_numpy(buf).to(engine.device)
    nw = engine.n_words(m)
    planes = engine.alloc_planes(n, nw)
    planes.fill_(0x5A5A5A5A)                                     # poison: pack must write every word
    engine.pack_codes(d, n, planes, nw)
    return planes, nw


@pytest.mark.parametrize("n,m,miss,ld", [
    (2, 1, 0.0, None), (3, 31, 0.0, None), (3, 32, 0.5, None), (5, 33, 0.1, 8), (50, 1000, 0.0, 52),
    (64, 64, 0.05, None), (65, 100, 0.05, 68), (67, 257, 1.0, None), (130, 4099, 0.02, 132),
    (200, 3, 0.0, 201),
])
def test_pack_planes_bit_exact(engine, n, m, miss, ld):
    rng = np.random.default_rng(n * 1000 + m)
    codes = _random_codes(rng, n, m, miss)
    planes, nw = _pack(engine, codes, ld)
    want = oracle.pack_planes(np.ascontiguousarray(codes.T))
    got = planes.cpu().numpy().view(np.uint32)
    assert np.array_equal(got, want)
    back = engine.unpack_codes(planes, n, nw, m).cpu().numpy()
    assert np.array_equal(back, codes.T)
    # canonical 2-bit sample-major matrix of SURVEY §8(c).1 from the unpacked codes
    assert np.array_equal(oracle.pack_canonical(back), oracle.pack_canonical(np.ascontiguousarray(codes.T)))


def test_pack_with_row_gather_and_word_offset(engine):
    rng = np.random.default_rng(5)
    n, rows_total = 37, 500
    codes = _random_codes(rng, n, rows_total, 0.1)
    keep = np.sort(rng.choice(rows_total, size=173, replace=False)).astype(np.int64)
    d = torch.from_numpy(codes).to(engine.device)
    nw_total = 10
    planes = engine.alloc_planes(n, nw_total)
    planes.fill_(-1)
    engine.pack_codes(d, n, planes, nw_total, word0=2, rows=torch.from_numpy(keep).to(engine.device))
    sel = np.ascontiguousarray(codes[keep].T)
    want = oracle.pack_planes(sel, n_words=6)
    got = planes.cpu().numpy().view(np.uint32).reshape(-1, nw_total, 2, 64)
    assert np.array_equal(got[:, 2:8].reshape(-1), want)
    assert (got[:, :2] == 0xFFFFFFFF).all() and (got[:, 8:] == 0xFFFFFFFF).all()


@pytest.mark.parametrize("variant", [1, 2, 3])
@pytest.mark.parametrize("n,m,miss", [
    (2, 10, 0.0), (3, 100, 0.3), (50, 4000, 0.0), (64, 2048, 0.05), (65, 1025, 0.05), (129, 5000, 0.0),
    (130, 3333, 0.5), (70, 40000, 0.001), (200, 70, 1.0),
])
def test_pair_counts_bit_exact(engine, n, m, miss, variant):
    rng = np.random.default_rng(n * 7 + m)
    codes = _random_codes(rng, n, m, miss)
    if n >= 50:
        codes[:, 3] = 3                                           # an all-missing sample
        codes[:, 5] = codes[:, 4]                                 # a duplicate pair
    planes, nw = _pack(engine, codes)
    counts = engine.alloc_counts(n)
    engine.pair_counts(planes, n, nw, counts, variant=variant)
    want = oracle.pair_counts(np.ascontiguousarray(codes.T))
    assert np.array_equal(counts.cpu().numpy().astype(np.int64), want)


def test_pair_counts_word_ranges_accumulate(engine):
    """Counting word ranges separately and accumulating == one pass (the multi-GPU identity)."""
    rng = np.random.default_rng(11)
    n, m = 90, 6400
    codes = _random_codes(rng, n, m, 0.03)
    planes, nw = _pack(engine, codes)
    whole = engine.alloc_counts(n)
    engine.pair_counts(planes, n, nw, whole)
    parts = engine.alloc_counts(n)
    for b, e in ((0, 7), (7, 64), (64, 65), (65, nw)):
        engine.pair_counts(planes, n, nw, parts, word_begin=b, word_end=e)
    assert torch.equal(whole, parts)


@pytest.mark.parametrize("metric", ["ibs", "p"])
def test_normalise_bit_exact(engine, metric):
    rng = np.random.default_rng(3)
    n, m = 40, 900
    codes = _random_codes(rng, n, m, 0.4)
    codes[:, 7] = 3
    want_counts = oracle.pair_counts(np.ascontiguousarray(codes.T))
    counts = torch.from_numpy(want_counts.astype(np.int32)).to(engine.device)
    dist, nz = engine.normalise(counts, metric)
    want, want_nz = oracle.normalise(want_counts, metric)
    assert np.array_equal(dist.cpu().numpy(), want)              # within 1e-12 relative: in fact bit-exact
    assert int(nz.item()) == want_nz and want_nz > 0


def _nj_counters(engine):
    """[rescans, list evaluations, iterations, matrices joined] of the bounded-scan kernel since the last reading."""
    import ctypes as C
    from phyloforge_b200._lib import check
    out = (C.c_ulonglong * 8)()
    check(engine.lib.pf_nj_profile(0, out))
    return list(out)[4:8]


def _check_nj(engine, dist, ld=None, expect_bounded=None):
    """All three NJ kernels through the diagnostic switch: the bounded scan (8: at every size it can take; by default
    it runs from 1,536 nodes on), the fast full-scan kernel (4) and the general one (2). expect_bounded: whether the
    bounded-scan kernel must have joined the matrix itself (True), must have handed it over (False), or either."""
    from phyloforge_b200._lib import check
    n = dist.shape[0]
    want_m, want_l = oracle.nj(dist)
    names = [f"t{i}" for i in range(n)]
    want_nw = merges_to_newick(want_m, want_l, names)
    for force_general in (8, 4, 2):
        check(engine.lib.pf_nj_profile(force_general, None))
        try:
            if ld is None:
                d = torch.from_numpy(dist).to(engine.device)
            else:                                                  # rows with a wider leading dimension
                buf = torch.full((n, ld), float("nan"), dtype=torch.float64, device=engine.device)
                buf[:, :n] = torch.from_numpy(dist).to(engine.device)
                d = buf[:, :n]
            merge, length = engine.nj(d)
            torch.cuda.synchronize()
        finally:
            joined = _nj_counters(engine)[3]
        if force_general == 8 and expect_bounded is not None and n >= 8:
            assert joined == (1 if expect_bounded else 0), "bounded-scan kernel: joined %d" % joined
        if force_general != 8:
            assert joined == 0
        got_nw = merges_to_newick(merge.cpu().numpy(), length.cpu().numpy(), names)
        rf, dl = treecmp.compare(got_nw, want_nw)
        which = {8: "bounded scan", 4: "fast kernel", 2: "general kernel"}[force_general]
        assert rf == 0, which                                        # Robinson-Foulds distance 0
        assert dl <= 1e-9, which                                     # branch lengths within 1e-9
        assert np.array_equal(merge.cpu().numpy(), want_m), which    # in fact the same join order
        assert np.array_equal(length.cpu().numpy(), want_l), which   # and bit-identical lengths


@pytest.mark.parametrize("n", [3, 4, 5, 6, 33, 64, 257, 513, 700, 1500])
def test_nj_random_distances(engine, n):
    rng = np.random.default_rng(n)
    a = rng.random((n, n))
    dist = (a + a.T) / 2
    np.fill_diagonal(dist, 0.0)
    _check_nj(engine, dist, expect_bounded=True)
    if n in (5, 257):
        _check_nj(engine, dist, ld=n + 3)


@pytest.mark.parametrize("k,n,mode", [(1, 30, 0), (5, 100, 0), (20, 40, 0), (3, 700, 0), (5, 100, 8), (11, 300, 8),
                                      (3, 1600, 0)])
def test_nj_batch_matches_oracle_tree_by_tree(engine, k, n, mode):
    """pf_nj_batch: several matrices joined concurrently - by groups of CTAs of one cooperative launch (full scan) or
    by one cluster each (bounded scan: from 1,536 nodes on, or forced for small ones with mode 8)."""
    from phyloforge_b200._lib import check
    rng = np.random.default_rng(k * 1000 + n)
    mats = []
    for _ in range(k):
        a = rng.random((n, n))
        d = (a + a.T) / 2
        np.fill_diagonal(d, 0.0)
        mats.append(d)
    check(engine.lib.pf_nj_profile(mode, None))
    try:
        merge, length = engine.nj_batch(torch.from_numpy(np.stack(mats)).to(engine.device))
        torch.cuda.synchronize()
    finally:
        joined = _nj_counters(engine)[3]
    assert joined == (k if (mode == 8 or n >= 1536) and n >= 8 else 0)
    for j in range(k):
        want_m, want_l = oracle.nj(mats[j])
        assert np.array_equal(merge[j].cpu().numpy(), want_m), j
        assert np.array_equal(length[j].cpu().numpy(), want_l), j


def test_nj_many_exact_ties(engine):
    """Rational distances with a common denominator: Q ties are real and must break the same way."""
    rng = np.random.default_rng(99)
    for n in (120, 600):
        a = rng.integers(0, 6, size=(n, n)).astype(np.float64) / 8.0
        dist = np.triu(a, 1)
        dist = dist + dist.T
        _check_nj(engine, dist)
    _check_nj(engine, np.ones((50, 50)) - np.eye(50))            # star: everything ties
    # ... across scan segments and CTAs; no bound can prune here: the bounded scan gives up and hands over
    _check_nj(engine, np.ones((1030, 1030)) - np.eye(1030), expect_bounded=False)


def test_nj_ultrametric_and_caterpillar_trees(engine):
    """Tree-like inputs: the node merged last is joined again at once (caterpillar), so the pending
    update of the fast kernel is read back through shared memory every iteration. The additive
    matrix is not exactly symmetric in floating point ((x + h_i) + h_j vs (x + h_j) + h_i): the
    kernels read the same triangle entries as the oracle, so the results are still bit-identical."""
    n = 400
    pos = np.cumsum(np.arange(1, n + 1, dtype=np.float64) * 0.01)
    height = np.linspace(0.5, 1.5, n)
    dist = np.abs(pos[:, None] - pos[None, :]) + height[:, None] + height[None, :]   # additive (caterpillar) metric
    np.fill_diagonal(dist, 0.0)
    _check_nj(engine, dist)
    rng = np.random.default_rng(1)
    perm = rng.permutation(n)
    _check_nj(engine, dist[np.ix_(perm, perm)])
    levels = np.array([[bin(i ^ j).count("1") and (i ^ j).bit_length() for j in range(256)] for i in range(256)],
                      dtype=np.float64)                                              # balanced ultrametric tree
    _check_nj(engine, levels)


def test_nj_non_finite_input_does_not_hang(engine):
    """NaN / inf distances: no candidate ever wins, slots 0 and 1 are joined (as the oracle does)."""
    n = 40
    dist = np.full((n, n), np.nan)
    np.fill_diagonal(dist, 0.0)
    d = torch.from_numpy(dist).to(engine.device)
    merge, _ = engine.nj(d)
    want_m, _ = oracle.nj(dist)
    assert np.array_equal(merge.cpu().numpy(), want_m)


def test_nj_from_genotypes(engine):
    n, m = 300, 20000
    codes = oracle.synth_codes(n, 0, m, seed=1, miss_ppm=50000)
    counts = oracle.pair_counts(np.ascontiguousarray(codes.T), method="bits")
    dist, _ = oracle.normalise(counts, "ibs")
    _check_nj(engine, dist, expect_bounded=True)


def test_nj_bounded_scan_hands_over_what_it_cannot_join(engine):
    """The bounded-scan kernel reads a distance from the row of either node, so it takes symmetric matrices only; an
    asymmetric input goes to the full-scan kernel (which reads the oracle's triangle) - same result either way."""
    rng = np.random.default_rng(5)
    n = 200
    a = rng.random((n, n))
    np.fill_diagonal(a, 0.0)
    _check_nj(engine, a, expect_bounded=False)                      # asymmetric
    sym = (a + a.T) / 2
    dup = sym.copy()
    dup[:, 50] = dup[:, 51]                                           # two identical samples ...
    dup[50, :] = dup[51, :]
    dup[50, 51] = dup[51, 50] = 0.0                                   # ... at distance 0
    dup[50, 50] = 0.0
    _check_nj(engine, dup)


def test_nj_bounded_scan_every_list_width(engine):
    """The list width (64, 32, 16 or 8 tracked columns per row) follows from shared memory; n = 3,600 and 6,500 take
    the narrower ones (the widest runs in every other test). Checked against the full-scan kernel."""
    from phyloforge_b200._lib import check
    for n in (3600, 6500):
        a = torch.rand((n, n), dtype=torch.float64, device=engine.device)
        d = (a + a.T) / 2
        d.fill_diagonal_(0)
        m0, l0 = engine.nj(d)
        torch.cuda.synchronize()
        assert _nj_counters(engine)[3] == 1
        check(engine.lib.pf_nj_profile(4, None))
        try:
            m1, l1 = engine.nj(d)
            torch.cuda.synchronize()
        finally:
            check(engine.lib.pf_nj_profile(0, None))
        assert torch.equal(m0, m1) and torch.equal(l0, l1), n


def test_synth_generator_matches_oracle_and_is_shard_invariant(engine):
    n = 123
    a = engine.synth_codes(n, 1000, 700, seed=42, miss_ppm=100000)[:, :n].cpu().numpy()
    assert np.array_equal(a, oracle.synth_codes(n, 1000, 700, seed=42, miss_ppm=100000))
    b = engine.synth_codes(n, 1300, 400, seed=42, miss_ppm=100000)[:, :n].cpu().numpy()
    assert np.array_equal(a[300:], b)
    big = engine.synth_codes(n, (1 << 33) + 5, 64, seed=42)[:, :n].cpu().numpy()   # 64-bit site index
    assert np.array_equal(big, oracle.synth_codes(n, (1 << 33) + 5, 64, seed=42))


def test_transpose_u8(engine):
    rng = np.random.default_rng(8)
    a = rng.integers(0, 255, size=(301, 77), dtype=np.uint8)
    rows = np.array([5, 0, 300, 17, 17, 250], dtype=np.int64)
    d = torch.from_numpy(a).to(engine.device)
    out = engine.transpose_u8(d, 77, rows=torch.from_numpy(rows).to(engine.device))
    assert np.array_equal(out.cpu().numpy(), a[rows].T)
    out2 = engine.transpose_u8(d, 77)
    assert np.array_equal(out2.cpu().numpy(), a.T)


def test_end_to_end_tree_matches_oracle(engine):
    n, m = 150, 30000
    names = [f"id{i}" for i in range(n)]
    codes = engine.synth_codes(n, 0, m, seed=3, miss_ppm=20000)
    nw = engine.n_words(m)
    planes = engine.alloc_planes(n, nw)
    engine.pack_codes(codes, n, planes, nw)
    for metric in ("ibs", "p"):
        res = engine.tree_from_planes(planes, n, nw, names, metric=metric)
        sm = np.ascontiguousarray(oracle.synth_codes(n, 0, m, seed=3, miss_ppm=20000).T)
        counts = oracle.pair_counts(sm, method="bits")
        dist, _ = oracle.normalise(counts, metric)
        merge, length = oracle.nj(dist)
        rf, dl = treecmp.compare(res.newick, merges_to_newick(merge, length, names))
        assert (rf, dl) == (0, 0.0)


def test_host_buffer_path_matches_resident_path(engine):
    """Engine.tree_from_host_codes (chunked H2D overlapped with pack + counts) == resident path."""
    n, m = 100, 10000
    names = [f"h{i}" for i in range(n)]
    codes = engine.synth_codes(n, 0, m, seed=5, miss_ppm=10000)
    host = codes.cpu().pin_memory()
    nw = engine.n_words(m)
    planes = engine.alloc_planes(n, nw)
    engine.pack_codes(codes, n, planes, nw)
    want = engine.tree_from_planes(planes, n, nw, names)
    for chunk in (1024, 4096, 1 << 20):
        got = engine.tree_from_host_codes(host, n, names, chunk_sites=chunk)
        assert torch.equal(got.counts, want.counts) and got.newick == want.newick
        assert np.array_equal(got.dist_host, want.dist.cpu().numpy()) and got.d2h_bytes > 0


@pytest.fixture(params=[(2, 0, False), (2, 1, False), (2, 2, False), (2, 3, False), (2, 4, False), (2, 5, False),
                        (2, 0, True), (1, 0, False)],
                ids=["pair", "pair-j240-regs", "pair-j240-tma", "pair-j208-regs", "pair-bulk-copies", "pair-j192-7steps",
                     "pair-strict", "single-cta"])
def tc_impl(engine, request):
    """Every GEMM kernel behind pf_pair_counts_tc: the CTA-pair kernel (default geometry, the alternatives, and
    with cluster-scope release on the hand-over) and the single-CTA kernel."""
    impl, cfg, strict = request.param
    engine.set_tc_impl(impl, cfg, strict)
    yield request.param
    engine.set_tc_impl()


@pytest.mark.parametrize("n,k", [(32, 64), (208, 256), (240, 1024), (256, 128)])
def test_cta_pair_mma_instruction_is_exact(engine, n, k):
    """tcgen05.mma.cta_group::2 kind::mxf4 on one CTA pair (256 x n x k, E2M1 operands, operands written by
    generic-proxy stores in both CTAs, remote arrive, multicast commit) == integer dot products."""
    from phyloforge_b200._lib import check
    rng = np.random.default_rng(n * 1000 + k)
    vals = np.array([-6, -4, -3, -2, -1, 0, 1, 2, 3, 4, 6], dtype=np.int8) if k <= 128 else np.array([-1, 0, 1], dtype=np.int8)
    a = rng.choice(vals, size=(256, k))
    b = rng.choice(vals, size=(n, k))
    da, db = torch.from_numpy(a).to(engine.device), torch.from_numpy(b).to(engine.device)
    dc = torch.full((256, n), -7.0, dtype=torch.float32, device=engine.device)
    _lib.check_probe(_lib.load_probes().pf_selftest_umma_fp4_2cta(da.data_ptr(), db.data_ptr(), dc.data_ptr(), n, k,
                                               torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert np.array_equal(dc.cpu().numpy().astype(np.int64), a.astype(np.int64) @ b.astype(np.int64).T)


@pytest.mark.parametrize("n,m", [(2, 40), (130, 5000), (256, 4096), (257, 4100), (300, 70000), (700, 20000)])
def test_pair_counts_tensor_core_variant_bit_exact(engine, tc_impl, n, m):
    """tcgen05 int8 indicator GEMM == oracle on data without mixed missingness (incl. sites missing in
    every sample and site padding); with mixed missingness it declines and the popcount kernel runs."""
    rng = np.random.default_rng(n + m)
    codes = rng.integers(0, 3, size=(m, n), dtype=np.uint8)
    codes[rng.random(m) < 0.01] = 3                               # whole sites missing: still uniform
    planes, nw = _pack(engine, codes)
    counts = engine.alloc_counts(n)
    ran = engine.pair_counts(planes, n, nw, counts, variant=3)
    assert ran == 3
    want = oracle.pair_counts(np.ascontiguousarray(codes.T), method="bits")
    assert np.array_equal(counts.cpu().numpy().astype(np.int64), want)
    # accumulate semantics + word ranges, as the chunked / sharded callers use them
    parts = engine.alloc_counts(n)
    cut = max(1, nw // 3)
    assert engine.pair_counts(planes, n, nw, parts, word_begin=0, word_end=cut, variant=3) == 3
    assert engine.pair_counts(planes, n, nw, parts, word_begin=cut, word_end=nw, variant=3) == 3
    assert torch.equal(parts, counts)
    assert engine.tc_launches == 1
    codes[5, 1] = 3                                               # one genotype missing in one sample
    planes, nw = _pack(engine, codes)
    counts = engine.alloc_counts(n)
    assert engine.pair_counts(planes, n, nw, counts, variant=3, tc_general=False) == 2    # declines -> popcount
    want = oracle.pair_counts(np.ascontiguousarray(codes.T), method="bits")
    assert np.array_equal(counts.cpu().numpy().astype(np.int64), want)
    counts.zero_()
    assert engine.pair_counts(planes, n, nw, counts, variant=3) == 3 and engine.tc_launches == 2
    assert np.array_equal(counts.cpu().numpy().astype(np.int64), want)


@pytest.mark.parametrize("n,m,miss", [(2, 100, 0.3), (64, 32, 0.5), (129, 700, 0.1), (193, 3000, 0.02), (257, 5000, 0.2),
                                      (385, 2048, 0.9), (1000, 40000, 0.05)])
def test_pair_counts_tensor_core_general_missingness_bit_exact(engine, tc_impl, n, m, miss):
    """Genotypes missing in some samples only: the XX/HH launch plus the TT/VV launch == oracle,
    also over word ranges and when accumulated on top of existing counts."""
    rng = np.random.default_rng(n * 7 + m)
    codes = _random_codes(rng, n, m, miss)
    codes[rng.random(m) < 0.01] = 3
    planes, nw = _pack(engine, codes)
    want = oracle.pair_counts(np.ascontiguousarray(codes.T), method="bits")
    counts = engine.alloc_counts(n)
    assert engine.pair_counts(planes, n, nw, counts, variant=3) == 3 and engine.tc_launches == 2
    assert np.array_equal(counts.cpu().numpy().astype(np.int64), want)
    parts = engine.alloc_counts(n)
    cut = max(1, nw // 2)
    engine.pair_counts(planes, n, nw, parts, word_begin=0, word_end=cut, variant=3)
    engine.pair_counts(planes, n, nw, parts, word_begin=cut, word_end=nw, variant=2)     # mixed with the popcount kernel
    assert torch.equal(parts, counts)


@pytest.mark.parametrize("n,m", [(2, 33), (16, 64), (17, 100), (63, 257), (64, 32), (130, 1000), (500, 4099)])
def test_pack_2bit_rows_bit_exact(engine, n, m):
    """Site-major 2-bit host format (and the PLINK .bed code table) -> the same planes as the byte packer."""
    rng = np.random.default_rng(n * 31 + m)
    codes = _random_codes(rng, n, m, 0.1)
    d = torch.from_numpy(codes).to(engine.device)
    nw = engine.n_words(m)
    want = oracle.pack_planes(np.ascontiguousarray(codes.T))
    rows2 = engine.to_2bit_rows(d, n)
    planes = engine.alloc_planes(n, nw)
    planes.fill_(0x5A5A5A5A)
    engine.pack_codes_2bit(rows2, n, planes, nw)
    assert np.array_equal(planes.cpu().numpy().view(np.uint32), want)
    # PLINK .bed codes: 00 hom A1, 01 missing, 10 het, 11 hom A2
    lut = torch.tensor([0b00, 0b10, 0b11, 0b01], dtype=torch.uint8, device=engine.device)
    bed = engine.to_2bit_rows(lut[d.long()], n)
    pad = (-n) % 16
    if pad:                                   # to_2bit_rows pads with our missing code 3 = 0b11; .bed pads are ignored
        pass
    planes.fill_(0x5A5A5A5A)
    engine.pack_codes_2bit(bed, n, planes, nw, plink=True)
    assert np.array_equal(planes.cpu().numpy().view(np.uint32), want)


def test_host_2bit_path_matches_byte_path(engine):
    n, m = 260, 20000
    names = [f"h{i}" for i in range(n)]
    codes = engine.synth_codes(n, 0, m, seed=6)
    host_u8 = codes.cpu().pin_memory()
    host_2b = engine.to_2bit_rows(codes, n).cpu().pin_memory()
    a = engine.tree_from_host_codes(host_u8, n, names, chunk_sites=4096)
    b = engine.tree_from_host_codes(host_2b, n, names, chunk_sites=4096, packed2=True)
    assert torch.equal(a.counts, b.counts) and a.newick == b.newick


def test_kernels_do_not_write_outside_their_buffers(engine):
    """compute-sanitizer is closed on this GPU pool, so out-of-bounds stores are looked for with
    canaries: every output buffer is over-allocated and the padding must stay untouched."""
    from phyloforge_b200._lib import check
    n, m = 203, 5000
    rng = np.random.default_rng(77)
    codes = _random_codes(rng, n, m, 0.0)
    d = torch.from_numpy(codes).to(engine.device)
    nw = engine.n_words(m)
    st = torch.cuda.current_stream().cuda_stream
    plen = int(engine.lib.pf_planes_len(n, nw))
    guard = 4096
    pbuf = torch.full((plen + 2 * guard,), 0x13572468, dtype=torch.int32, device=engine.device)
    planes = pbuf[guard:guard + plen]
    engine.pack_codes(d, n, planes, nw)
    assert bool((pbuf[:guard] == 0x13572468).all()) and bool((pbuf[guard + plen:] == 0x13572468).all())
    engine.pack_codes_2bit(engine.to_2bit_rows(d, n), n, planes, nw)
    assert bool((pbuf[:guard] == 0x13572468).all()) and bool((pbuf[guard + plen:] == 0x13572468).all())
    ld = n + 5                                                    # counts with a wider leading dimension
    for variant in (1, 2, 3):
        cbuf = torch.zeros((3 * n * ld + 2 * guard,), dtype=torch.int32, device=engine.device)
        cbuf[:guard] = -99
        cbuf[guard + 3 * n * ld:] = -99
        cview = cbuf[guard:guard + 3 * n * ld]
        if variant == 3:
            work = torch.empty(int(engine.lib.pf_pair_counts_tc_work_bytes(n, nw)) + 2 * guard, dtype=torch.uint8,
                               device=engine.device)
            work.fill_(0xA5)
            import ctypes as C
            used = C.c_int(0)
            check(engine.lib.pf_pair_counts_tc(planes.data_ptr(), n, nw, 0, nw, cview.data_ptr(), ld,
                                               work[guard:].data_ptr(), work.numel() - 2 * guard, 1, C.byref(used), st))
            assert used.value == 1
            torch.cuda.synchronize()
            assert bool((work[:guard] == 0xA5).all()) and bool((work[-guard:] == 0xA5).all())
        else:
            check(engine.lib.pf_pair_counts(planes.data_ptr(), n, nw, 0, nw, cview.data_ptr(), ld, variant, st))
        torch.cuda.synchronize()
        assert bool((cbuf[:guard] == -99).all()) and bool((cbuf[guard + 3 * n * ld:] == -99).all())
        got = cview.view(3, n, ld)
        assert bool((got[:, :, n:] == 0).all())                   # padding columns of the counts are never touched
        want = oracle.pair_counts(np.ascontiguousarray(codes.T), method="bits")
        assert np.array_equal(got[:, :, :n].cpu().numpy().astype(np.int64), want)
    # NJ: workspace and outputs
    dist = torch.rand((n, n), dtype=torch.float64, device=engine.device)
    dist = (dist + dist.T) / 2
    dist.fill_diagonal_(0)
    wbytes = int(engine.lib.pf_nj_work_bytes(n))
    work = torch.full((wbytes + 2 * guard,), 0x5C, dtype=torch.uint8, device=engine.device)
    mbuf = torch.full((3 * (n - 2) + 64,), -7, dtype=torch.int32, device=engine.device)
    lbuf = torch.full((3 * (n - 2) + 64,), -7.0, dtype=torch.float64, device=engine.device)
    dcopy = dist.clone()
    check(engine.lib.pf_nj(dcopy.data_ptr(), n, n, mbuf.data_ptr(), lbuf.data_ptr(), work[guard:].data_ptr(), st))
    torch.cuda.synchronize()
    assert bool((work[:guard] == 0x5C).all()) and bool((work[guard + wbytes:] == 0x5C).all())
    assert bool((mbuf[3 * (n - 2):] == -7).all()) and bool((lbuf[3 * (n - 2):] == -7.0).all())
