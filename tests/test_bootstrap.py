"""Block bootstrap: host logic and oracle on the CPU, kernels against the oracle on the GPU."""
import numpy as np
import pytest

from oracle import cpu as oracle, treecmp
from phyloforge_b200.bootstrap import split_keys, split_support
from phyloforge_b200.newick import merges_to_newick
from phyloforge_b200.sharding import shard_words


def _exact_support(main_m, main_l, rep_ms, rep_ls, names):
    """Support by comparing bipartitions as leaf sets (no hashing)."""
    n = len(names)
    main = treecmp.parse_newick(merges_to_newick(main_m, main_l, names))
    reps = [set(k for k in treecmp.splits(treecmp.parse_newick(merges_to_newick(m, l, names))) if 1 < len(k) < n - 1)
            for m, l in zip(rep_ms, rep_ls)]
    # node n+t of the join list <-> the leaf set below it
    below = {i: frozenset([names[i]]) for i in range(n)}
    ref = sorted(names)[0]
    out = []
    full = frozenset(names)
    for t in range(n - 3):
        s = below[int(main_m[t][0])] | below[int(main_m[t][1])]
        below[n + t] = s
        key = full - s if ref in s else s
        out.append(100.0 * sum(key in r for r in reps) / len(reps))
    assert set(k for k in treecmp.splits(main) if 1 < len(k) < n - 1) == \
        set((full - below[n + t]) if ref in below[n + t] else below[n + t] for t in range(n - 3))
    return np.array(out)


def test_bootstrap_weights_are_a_multinomial_draw_and_counter_based():
    w = oracle.bootstrap_weights(7, 0, 50, 64)
    assert w.shape == (50, 64) and (w.sum(1) == 64).all() and w.min() >= 0
    assert np.array_equal(oracle.bootstrap_weights(7, 20, 5, 64), w[20:25])       # replicate index is the counter
    assert not np.array_equal(oracle.bootstrap_weights(8, 0, 50, 64), w)
    frac_zero = (w == 0).mean()
    assert 0.30 < frac_zero < 0.44                                               # ~ 1/e of the blocks are left out


def test_split_support_matches_exact_bipartitions():
    rng = np.random.default_rng(5)
    n = 30
    names = [f"s{i:02d}" for i in range(n)]
    a = rng.random((n, n))
    d = (a + a.T) / 2
    np.fill_diagonal(d, 0)
    main_m, main_l = oracle.nj(d)
    rep_ms, rep_ls = [], []
    for r in range(25):
        noise = rng.random((n, n)) * (0.02 if r % 2 else 0.3)
        e = d + (noise + noise.T) / 2
        np.fill_diagonal(e, 0)
        m, l = oracle.nj(e)
        rep_ms.append(m)
        rep_ls.append(l)
    got = split_support(main_m, np.stack(rep_ms), n)
    want = _exact_support(main_m, main_l, rep_ms, rep_ls, names)
    assert np.array_equal(got, want)
    assert got.max() <= 100.0 and got.min() >= 0.0 and len(np.unique(got)) > 2
    assert (split_support(main_m, np.stack([main_m] * 3), n) == 100.0).all()
    # keys are a function of the split only: relabelled join order of the same tree gives the same key set
    assert set(split_keys(main_m, n)[0]) == set(split_keys(main_m.copy(), n)[0])


def test_newick_support_labels_round_trip():
    rng = np.random.default_rng(2)
    n = 9
    names = [f"t{i}" for i in range(n)]
    a = rng.random((n, n))
    d = (a + a.T) / 2
    np.fill_diagonal(d, 0)
    m, l = oracle.nj(d)
    support = np.arange(n - 3, dtype=np.float64) * 12.5
    text = merges_to_newick(m, l, names, support=support)
    plain = merges_to_newick(m, l, names)
    assert treecmp.compare(text, plain) == (0, 0.0)            # labels do not change topology or lengths
    labels = []
    st = [treecmp.parse_newick(text)]
    while st:
        v = st.pop()
        if v.children:
            st.extend(v.children)
            if v.name not in (None, ""):
                labels.append(float(v.name))
    assert sorted(labels) == sorted(support.tolist())


def test_oracle_bootstrap_counts_of_unit_weights_is_the_total():
    rng = np.random.default_rng(3)
    n, m, B = 12, 2000, 8
    codes = rng.integers(0, 4, size=(n, m), dtype=np.uint8)
    nw = (m + 31) // 32
    blocks = []
    for b in range(B):
        wb, we = shard_words(nw, B, b)
        blocks.append(oracle.pair_counts(np.ascontiguousarray(codes[:, wb * 32:min(m, we * 32)])))
    blocks = np.stack(blocks)
    total = oracle.bootstrap_counts(blocks, np.ones((1, B), dtype=np.int32))[0]
    assert np.array_equal(total, oracle.pair_counts(codes))


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_gpu_bootstrap_weights_and_counts_match_oracle(engine):
    import torch
    for seed, r0, R, B in ((1, 0, 8, 16), (99, 13, 5, 257), (5, 1000, 1, 1)):
        w = engine.bootstrap_weights(seed, r0, R, B)
        assert np.array_equal(w.cpu().numpy(), oracle.bootstrap_weights(seed, r0, R, B))
    rng = np.random.default_rng(0)
    for n, B, R in ((5, 3, 1), (7, 9, 8), (33, 20, 3)):            # 3 n^2 not a multiple of 4: padded path
        blocks = rng.integers(-1000, 1000000, size=(B, 3, n, n), dtype=np.int32)
        w = oracle.bootstrap_weights(3, 0, R, B)
        got = engine.bootstrap_counts(torch.from_numpy(blocks).to(engine.device), torch.from_numpy(w).to(engine.device))
        assert np.array_equal(got.cpu().numpy().astype(np.int64), oracle.bootstrap_counts(blocks, w))


@pytest.mark.gpu
@pytest.mark.parametrize("n,m,miss,B,R", [(40, 30000, 0, 16, 20), (210, 20000, 50000, 12, 9)])
def test_gpu_bootstrap_tree_matches_oracle_replicate_by_replicate(engine, n, m, miss, B, R):
    import torch
    names = [f"s{i:03d}" for i in range(n)]
    codes = engine.synth_codes(n, 0, m, seed=11, miss_ppm=miss)
    nw = engine.n_words(m)
    planes = engine.alloc_planes(n, nw)
    engine.pack_codes(codes, n, planes, nw)
    main, support, rep_merges = engine.bootstrap_tree(planes, n, nw, names, replicates=R, seed=4, n_blocks=B)
    torch.cuda.synchronize()
    # oracle: the same blocks, draws, weighted sums, distances and joins
    sm = np.ascontiguousarray(oracle.synth_codes(n, 0, m, seed=11, miss_ppm=miss).T)
    blocks = []
    for b in range(B):
        wb, we = shard_words(nw, B, b)
        blocks.append(oracle.pair_counts(np.ascontiguousarray(sm[:, wb * 32:min(m, we * 32)]), method="bits"))
    blocks = np.stack(blocks)
    total = blocks.sum(0)
    assert np.array_equal(main.counts.cpu().numpy().astype(np.int64), total)
    dist, _ = oracle.normalise(total, "ibs")
    main_m, main_l = oracle.nj(dist)
    assert np.array_equal(main.merge, main_m) and np.array_equal(main.length, main_l)
    w = oracle.bootstrap_weights(4, 0, R, B)
    rep = oracle.bootstrap_counts(blocks, w)
    rep_ms, rep_ls = [], []
    for r in range(R):
        d, _ = oracle.normalise(rep[r], "ibs")
        mm, ll = oracle.nj(d)
        assert np.array_equal(rep_merges[r], mm), f"replicate {r}"
        rep_ms.append(mm)
        rep_ls.append(ll)
    want = _exact_support(main_m, main_l, rep_ms, rep_ls, names)
    assert np.array_equal(support, want)
    assert treecmp.compare(main.newick, merges_to_newick(main_m, main_l, names)) == (0, 0.0)
    assert 0.0 <= support.min() and support.max() <= 100.0


@pytest.mark.gpu
def test_gpu_bootstrap_replicates_in_groups_of_nine_eights_match_oracle(engine):
    """From 1,536 samples on the replicates are joined by the bounded-scan kernel, one cluster per matrix, in groups of
    up to 72 (nine eights): every replicate's join list equals the oracle's for the same draw, whatever group and
    position it was joined in (80 replicates = a group of 72 and a group of 8)."""
    import torch
    n, m, B, R = 1600, 6400, 8, 80
    names = [f"s{i:04d}" for i in range(n)]
    codes = engine.synth_codes(n, 0, m, seed=23)
    nw = engine.n_words(m)
    planes = engine.alloc_planes(n, nw)
    engine.pack_codes(codes, n, planes, nw)
    main, support, rep_merges = engine.bootstrap_tree(planes, n, nw, names, replicates=R, seed=9, n_blocks=B, metric="p")
    torch.cuda.synchronize()
    sm = np.ascontiguousarray(oracle.synth_codes(n, 0, m, seed=23).T)
    blocks = []
    for b in range(B):
        wb, we = shard_words(nw, B, b)
        blocks.append(oracle.pair_counts(np.ascontiguousarray(sm[:, wb * 32:min(m, we * 32)]), method="bits"))
    blocks = np.stack(blocks)
    dist, _ = oracle.normalise(blocks.sum(0), "p")
    main_m, main_l = oracle.nj(dist)
    assert np.array_equal(main.merge, main_m) and np.array_equal(main.length, main_l)
    w = oracle.bootstrap_weights(9, 0, R, B)
    for r in (0, 7, 8, 35, 63, 71, 72, 79):                       # first / last of an eight, of the group of 72, of the rest
        rep = oracle.bootstrap_counts(blocks, w[r:r + 1])[0]
        d, _ = oracle.normalise(rep, "p")
        mm, _ = oracle.nj(d)
        assert np.array_equal(rep_merges[r], mm), f"replicate {r}"
    assert len({rep_merges[r].tobytes() for r in range(R)}) > R // 2   # the replicates are different trees
    assert support.shape == (n - 3,) and 0.0 <= support.min() and support.max() <= 100.0


def _random_join_list(n, rng):
    live = list(range(n))
    m = np.zeros((n - 2, 3), dtype=np.int32)
    for t in range(n - 3):
        i = int(rng.integers(len(live)))
        a = live[i]
        live[i] = live[-1]
        live.pop()
        j = int(rng.integers(len(live)))
        b = live[j]
        live[j] = n + t
        m[t] = (a, b, 0)
    m[n - 3] = live
    return m


@pytest.mark.parametrize("n,R", [(4, 3), (5, 40), (12, 50), (97, 130), (1000, 37)])
def test_native_split_support_equals_numpy(n, R):
    """pf_bootstrap_support (host threads of the library) against bootstrap.split_support_py: replicates that are the
    main tree, relabelled main trees (some splits survive), random trees; any number of threads."""
    from phyloforge_b200.bootstrap import split_support, split_support_py
    rng = np.random.default_rng(n + R)
    main = _random_join_list(n, rng)
    reps = []
    for r in range(R):
        kind = r % 3
        if kind == 0:
            reps.append(main.copy())
        elif kind == 1:                                   # swap two leaves of the main tree
            m = main.copy()
            a, b = (int(x) for x in rng.choice(n, 2, replace=False))
            m[main == a], m[main == b] = b, a
            reps.append(m)
        else:
            reps.append(_random_join_list(n, rng))
    reps = np.stack(reps)
    want = split_support_py(main, reps, n)
    for threads in (1, 3, 0):
        got = split_support(main, reps, n, threads=threads)
        assert np.array_equal(got, want), threads
    assert want.max() <= 100.0 and want.min() >= 100.0 * ((R + 2) // 3) / R - 1e-9   # every third replicate is the main tree
    assert np.array_equal(split_support(main, reps[:0], n), np.zeros(n - 3))       # no replicates: zeros
    bad = reps.copy()
    bad[R // 2, 0, 0] = n + 5                                                       # a node that does not exist yet
    from phyloforge_b200._lib import PhyloForgeError
    with pytest.raises(PhyloForgeError):
        split_support(main, bad, n)
