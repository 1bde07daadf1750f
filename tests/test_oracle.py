"""The oracle itself: C restatement vs independent numpy restatements, known-answer NJ cases,
and size-independent properties. (PARITY UNPINNED for counts/distance/NJ: no treebest here.)"""
import numpy as np
import pytest

from oracle import cpu as oracle, treecmp
from phyloforge_b200.newick import merges_to_newick


def _rand_codes(rng, n, m, miss):
    c = rng.integers(0, 3, size=(n, m), dtype=np.uint8)
    c[rng.random((n, m)) < miss] = 3
    return c


@pytest.mark.parametrize("n,m,miss", [(2, 5, 0.0), (7, 100, 0.3), (33, 257, 0.1), (20, 64, 1.0)])
def test_counts_c_vs_numpy_vs_bits(n, m, miss):
    rng = np.random.default_rng(n + m)
    codes = _rand_codes(rng, n, m, miss)
    a = oracle.pair_counts(codes, method="bytes")
    assert np.array_equal(a, oracle.pair_counts_numpy(codes))
    assert np.array_equal(a, oracle.pair_counts(codes, method="bits"))
    V, D1, D2 = a
    assert (V == V.T).all() and (D1 == D1.T).all() and (D2 == D2.T).all()
    assert (np.diag(D1) == 0).all() and (D2 <= D1).all() and (D1 <= V).all()
    assert np.array_equal(np.diag(V), (codes != 3).sum(1))


def test_counts_known_answer():
    codes = np.array([[0, 1, 2, 3, 0, 2], [2, 1, 0, 0, 3, 2], [0, 0, 0, 0, 0, 0]], dtype=np.uint8)
    V, D1, D2 = oracle.pair_counts(codes)
    assert V[0, 1] == 4 and D1[0, 1] == 2 and D2[0, 1] == 2
    assert V[0, 2] == 5 and D1[0, 2] == 3 and D2[0, 2] == 2
    d_ibs, nz = oracle.normalise(oracle.pair_counts(codes), "ibs")
    d_p, _ = oracle.normalise(oracle.pair_counts(codes), "p")
    assert nz == 0 and d_ibs[0, 1] == (2 + 2) / 8 and d_p[0, 1] == 2 / 4 and d_ibs[0, 0] == 0


def test_sv_codes_ibs_equals_p():
    rng = np.random.default_rng(1)
    codes = rng.choice(np.array([0, 2, 3], dtype=np.uint8), size=(9, 300))
    c = oracle.pair_counts(codes)
    assert np.array_equal(c[1], c[2])                      # D2 == D1 when no het code exists
    assert np.array_equal(oracle.normalise(c, "ibs")[0], oracle.normalise(c, "p")[0])


def test_zero_valid_pairs_are_flagged():
    codes = np.array([[3, 3, 0], [0, 1, 3], [0, 0, 0]], dtype=np.uint8)
    d, nz = oracle.normalise(oracle.pair_counts(codes), "ibs")
    assert nz == 2 and d[0, 1] == 1.0 and d[1, 0] == 1.0


def test_pack_canonical_layout():
    codes = np.array([[0, 1, 2, 3, 1]], dtype=np.uint8)
    p = oracle.pack_canonical(codes)
    assert p.shape == (1, 128)
    assert p[0, 0] == (0 | 1 << 2 | 2 << 4 | 3 << 6) and p[0, 1] == (1 | 0xFC) and (p[0, 2:] == 0xFF).all()


def test_pack_planes_layout():
    rng = np.random.default_rng(2)
    codes = _rand_codes(rng, 70, 45, 0.2)
    planes = oracle.pack_planes(codes).reshape(2, 2, 2, 64)   # tiles, words, plane, lane
    for i, s in [(0, 0), (3, 31), (63, 32), (64, 44), (69, 7)]:
        lo = (planes[i // 64, s // 32, 0, i % 64] >> (s % 32)) & 1
        hi = (planes[i // 64, s // 32, 1, i % 64] >> (s % 32)) & 1
        assert lo | (hi << 1) == codes[i, s]
    assert (planes[1, :, :, 6:] == 0xFFFFFFFF).all()          # padding samples are code 3
    assert (planes[0, 1, :, 0] >> 13 == 0x7FFFF).all()        # padding sites are code 3


# ---------------------------------------------------------------- NJ known answers
WIKI = np.array([[0, 5, 9, 9, 8], [5, 0, 10, 10, 9], [9, 10, 0, 8, 7], [9, 10, 8, 0, 3], [8, 9, 7, 3, 0]], float)


def test_nj_textbook_five_taxa():
    """The worked example of the Wikipedia 'Neighbor joining' article (Saitou & Nei's method):
    a, b join first with branches 2 and 3; final tree ((a:2,b:3):3,c:4,(d:2,e:1):2)."""
    merge, length = oracle.nj(WIKI)
    assert tuple(merge[0][:2]) == (0, 1) and tuple(length[0][:2]) == (2.0, 3.0)
    got = merges_to_newick(merge, length, list("abcde"))
    rf, dl = treecmp.compare(got, "((a:2,b:3):3,c:4,(d:2,e:1):2);")
    assert rf == 0 and dl < 1e-12


@pytest.mark.parametrize("n", [3, 4, 6, 17, 40])
def test_nj_c_vs_python_restatement(n):
    rng = np.random.default_rng(n)
    a = rng.random((n, n))
    d = (a + a.T) / 2
    np.fill_diagonal(d, 0)
    m1, l1 = oracle.nj(d)
    m2, l2 = oracle.nj_numpy(d)
    assert np.array_equal(m1, m2) and np.array_equal(l1, l2)


def _random_tree_distances(rng, n):
    """Additive distances of a random unrooted binary tree (NJ must recover it exactly)."""
    names = [f"L{i}" for i in range(n)]
    nodes = [(nm, None) for nm in names]
    while len(nodes) > 3:
        i, j = sorted(rng.choice(len(nodes), 2, replace=False))
        a, b = nodes[i], nodes[j]
        la, lb = rng.uniform(0.01, 1.0, 2)
        new = (f"({a[0]}:{la:.17g},{b[0]}:{lb:.17g})", None)
        nodes = [x for k, x in enumerate(nodes) if k not in (i, j)] + [new]
    ls = rng.uniform(0.01, 1.0, 3)
    newick = "(" + ",".join(f"{x[0]}:{l:.17g}" for x, l in zip(nodes, ls)) + ");"
    tree = treecmp.parse_newick(newick)
    # path lengths via splits: d(x, y) = sum of branch lengths of splits separating x and y
    sp = treecmp.splits(tree)
    ref = sorted(names)[0]
    d = np.zeros((n, n))
    for side, ln in sp.items():
        inside = np.array([nm in side for nm in names])
        d += ln * (inside[:, None] != inside[None, :])
    return names, newick, d


@pytest.mark.parametrize("n", [5, 12, 60])
def test_nj_recovers_additive_tree(n):
    rng = np.random.default_rng(100 + n)
    names, truth, d = _random_tree_distances(rng, n)
    merge, length = oracle.nj(d)
    rf, dl = treecmp.compare(merges_to_newick(merge, length, names, fmt="%.17g"), truth)
    assert rf == 0 and dl < 1e-9


def test_nj_tie_break_is_lowest_slot_pair():
    d = np.ones((6, 6)) - np.eye(6)
    merge, _ = oracle.nj(d)
    assert tuple(merge[0][:2]) == (0, 1)


def test_synth_generator_properties():
    a = oracle.synth_codes(100, 0, 2000, seed=42, miss_ppm=100000)
    assert a.shape == (2000, 100) and a.max() == 3
    assert abs((a == 3).mean() - 0.1) < 0.01
    assert np.array_equal(a[500:900], oracle.synth_codes(100, 500, 400, seed=42, miss_ppm=100000))
    rows = np.array([3, 50, 99])
    assert np.array_equal(oracle.synth_rows(rows, 100, 0, 2000, seed=42, miss_ppm=100000), a[:, rows].T)
    b = oracle.synth_codes(100, 0, 2000, seed=43)
    assert (b != 3).all() and not np.array_equal(a, b)
    # population structure: samples of one block are closer to each other than to other blocks
    c = oracle.pair_counts(np.ascontiguousarray(oracle.synth_codes(100, 0, 4000, seed=1).T))
    d, _ = oracle.normalise(c, "ibs")
    assert d[0, 1:25].mean() < d[0, 25:].mean()


def test_treecmp_detects_differences():
    a = "((a:1,b:1):1,c:1,(d:1,e:1):1);"
    assert treecmp.compare(a, "((a:1,c:1):1,b:1,(d:1,e:1):1);")[0] == 2
    assert treecmp.compare(a, "((a:1,b:1.5):1,c:1,(d:1,e:1):1);") == (0, 0.5)
    assert treecmp.compare(a, "(c:1,(e:1,d:1):1,(b:1,a:1):1);") == (0, 0.0)
    assert treecmp.compare(a, "(((a:1,b:1):1,c:1):0.4,(d:1,e:1):0.6);") == (0, 0.0)   # rooted form


@pytest.mark.parametrize("n", [10, 80, 300])
def test_nj_incremental_row_sums_vs_from_scratch(n):
    """The oracle maintains the row sums incrementally (fixed operation order shared with the GPU
    kernel). On generic tie-free distances that gives the same topology as recomputing every sum
    from scratch each iteration, and branch lengths that agree far inside 1e-9."""
    rng = np.random.default_rng(n)
    a = rng.random((n, n))
    d = (a + a.T) / 2
    np.fill_diagonal(d, 0)
    names = [f"x{i}" for i in range(n)]
    m1, l1 = oracle.nj(d)
    m2, l2 = oracle.nj(d, from_scratch=True)
    rf, dl = treecmp.compare(merges_to_newick(m1, l1, names, fmt="%.17g"), merges_to_newick(m2, l2, names, fmt="%.17g"))
    assert rf == 0 and dl < 1e-11


def test_char_pair_counts_known_answer_and_agreement_with_genotype_counts():
    """oracle.char_pair_counts / char_mismatch_counts: a hand-checked case, and on biallelic characters they equal
    the genotype-code counts (D1 is the character mismatch on the reference's FASTA, D1 + D2 the
    allele-sharing distance |g_i - g_j|)."""
    chars = np.array([list(b"ACGN-Ta"), list(b"ACTNGTc"), list(b"MC-NGtA")], dtype=np.uint8)
    got = oracle.char_pair_counts(chars)
    assert got[0].tolist() == [[5, 5, 4], [5, 6, 5], [4, 5, 5]]
    assert got[1].tolist() == [[0, 2, 3], [2, 0, 3], [3, 3, 0]]      # (0,1): G/T, a/c; (0,2): A/M, T/t, a/A; (1,2): A/M, T/t, c/A
    assert got[2].tolist() == [[0, 1, 0], [1, 0, 1], [0, 1, 0]]      # no allele shared: G/T; c (C*) / A
    assert np.array_equal(oracle.char_mismatch_counts(chars), got[:2])
    rng = np.random.default_rng(4)
    codes = rng.integers(0, 4, size=(9, 300)).astype(np.uint8)          # REF = A, ALT = G, het = R, missing = N
    lut = np.frombuffer(b"ARGN", dtype=np.uint8)
    assert np.array_equal(oracle.char_pair_counts(lut[codes]), oracle.pair_counts(codes))
    with pytest.raises(ValueError):
        oracle.char_pair_counts(np.array([list(b"AX")], dtype=np.uint8))
