"""GPU counts / distances / NJ at the full size of the BASELINE configurations against the SPECIFIED oracle:
SURVEY.md §8(c).4 defines NJ with every row sum recomputed from scratch each iteration and lets the GPU update
them incrementally only if the trees agree (RF = 0, branch lengths within 1e-9) on all configurations. These
tests are that proof: C1, C3 and C4 against committed fixtures the CPU oracle produced at full size
(tests/golden/make_fullsize_fixtures.py: sha256 of the int32 count matrices and of the fp64 distance matrices,
join lists + branch lengths of both oracle forms), C2 through the SV converter, C5 with both oracle forms run
once on the GPU's own distance matrix (the oracle joins 10,000 nodes in tens of seconds)."""
import hashlib
import os
import time

import numpy as np
import pytest
import torch

from oracle import cpu as oracle, treecmp
from phyloforge_b200 import synth, vcf
from phyloforge_b200.newick import merges_to_newick

pytestmark = pytest.mark.gpu
FMT = "%.17g"


def _sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _counts(engine, n, m, miss, chunk_sites=1 << 21, seed=42):
    counts = engine.alloc_counts(n)
    nw = chunk_sites // 32
    planes = engine.alloc_planes(n, nw)
    for s0 in range(0, m, chunk_sites):
        k = min(chunk_sites, m - s0)
        codes = engine.synth_codes(n, s0, k, seed=seed, miss_ppm=miss)
        engine.pack_codes(codes, n, planes, nw, n_sites=k)
        engine.pair_counts(planes, n, nw, counts, word_end=(k + 31) // 32)
        del codes
    return counts


def _compare_with_both_oracles(merge, length, inc, fs, names, what):
    """GPU join list / branch lengths: bit-identical to the incremental oracle, and the same tree as the
    from-scratch oracle (RF = 0, branch lengths within 1e-9)."""
    assert np.array_equal(merge, inc[0]), f"{what}: join list differs from the incremental oracle"
    assert np.array_equal(length, inc[1]), f"{what}: branch lengths differ from the incremental oracle"
    gpu_nwk = merges_to_newick(merge, length, names, fmt=FMT)
    rf, dl = treecmp.compare(gpu_nwk, merges_to_newick(fs[0], fs[1], names, fmt=FMT))
    n_diff = int((np.asarray(merge) != np.asarray(fs[0])).any(axis=1).sum())
    print(f"{what}: vs from-scratch oracle RF = {rf}, max |branch length diff| = {dl:.3g}, "
          f"{n_diff} of {len(merge)} joins listed differently")
    assert rf == 0, f"{what}: topology differs from the from-scratch oracle (RF = {rf})"
    assert dl <= 1e-9, f"{what}: branch lengths differ from the from-scratch oracle by {dl}"


@pytest.mark.parametrize("cfg", ["c1", "c3", "c4"])
def test_full_size_counts_distances_and_nj_against_oracle_fixtures(engine, golden_dir, cfg):
    fx = np.load(os.path.join(golden_dir, f"fullsize_{cfg}.npz"))
    n, m, miss = int(fx["n"]), int(fx["m"]), int(fx["miss_ppm"])
    names = [f"s{i}" for i in range(n)]
    counts = _counts(engine, n, m, miss)
    assert _sha(counts.cpu().numpy()) == str(fx["counts_sha256"]), "int32 count matrices differ from the CPU oracle's"
    for metric in ("ibs", "p"):
        dist, _ = engine.normalise(counts, metric)
        assert _sha(dist.cpu().numpy()) == str(fx[f"dist_sha256_{metric}"]), f"{metric} distances differ"
        merge, length = engine.nj(dist)
        _compare_with_both_oracles(merge.cpu().numpy(), length.cpu().numpy(),
                                   (fx[f"merge_incremental_{metric}"], fx[f"length_incremental_{metric}"]),
                                   (fx[f"merge_from_scratch_{metric}"], fx[f"length_from_scratch_{metric}"]),
                                   names, f"{cfg} {n} x {m} {metric}")


def test_config2_sv_matrix_nj_against_from_scratch_oracle(engine, tmp_path):
    """C2 (500 samples x 100k SVs, 5 % missing) through the SV converter; both metrics coincide on SV columns."""
    n, m = 500, 100_000
    names = [f"V{i:03d}" for i in range(n)]
    codes = oracle.synth_codes(n, 0, m, seed=7, miss_ppm=50_000)
    text, svtype = synth.sv_vcf_bytes(codes, names)
    path = tmp_path / "c2.vcf"
    path.write_bytes(text)
    packed = vcf.load_sv_vcf(engine, str(path), keep_chars=False)
    res = engine.tree_from_planes(packed.planes, n, packed.n_words, names, metric="ibs", n_sites=packed.n_sites)
    want = oracle.pair_counts(synth.expected_sv_columns(codes, svtype), method="bits")
    assert np.array_equal(res.counts.cpu().numpy().astype(np.int64), want)
    dist, _ = oracle.normalise(want, "ibs")
    assert np.array_equal(res.dist.cpu().numpy(), dist)
    _compare_with_both_oracles(res.merge, res.length, oracle.nj(dist), oracle.nj(dist, from_scratch=True), names,
                               "c2 500 x 100k SV")


def test_config5_10000x5M_nj_against_the_oracle_not_the_gpu(engine):
    """C5 (10,000 x 5M, 10 % missing): the GPU's NJ on the GPU's distance matrix against BOTH oracle forms run
    on the same matrix on the host (not GPU-vs-GPU). The count matrices themselves are checked against the
    oracle on row subsets in test_gpu_fullsize.py."""
    n, m, miss = 10000, 5_000_000, 100_000
    names = [f"s{i}" for i in range(n)]
    counts = _counts(engine, n, m, miss, chunk_sites=1 << 20)
    dist, _ = engine.normalise(counts, "ibs")
    dist_h = dist.cpu().numpy()
    ref_dist, _ = oracle.normalise(counts.cpu().numpy().astype(np.int64), "ibs")
    assert np.array_equal(dist_h, ref_dist)
    del counts, ref_dist
    t0 = time.perf_counter()
    merge, length = engine.nj(dist)
    merge, length = merge.cpu().numpy(), length.cpu().numpy()
    t1 = time.perf_counter()
    inc = oracle.nj(dist_h)
    t2 = time.perf_counter()
    fs = oracle.nj(dist_h, from_scratch=True)
    t3 = time.perf_counter()
    print(f"c5 NJ at N = {n}: GPU {t1 - t0:.2f} s, oracle incremental {t2 - t1:.1f} s, from scratch {t3 - t2:.1f} s "
          f"({oracle.num_threads()} host threads)")
    _compare_with_both_oracles(merge, length, inc, fs, names, "c5 10000 x 5M ibs")
