"""The ALGORITHM of the bounded-scan NJ kernel (csrc/nj_lazy.cu) on the CPU: tools/nj_bounded_scan_model.c keeps the
kernel's per-row state (tracked columns, champion, the bounds R1 / R2 in the frame of a float snapshot, the drift
maximum) and takes its decisions sequentially; every join must be the join of oracle/pf_oracle.c:pfo_nj. This pins the
proof obligations of the design (nothing that is skipped could have won; ties break as the oracle breaks them) without
a GPU; the kernel itself is compared with the oracle in tests/test_gpu_kernels.py."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def model(tmp_path_factory):
    exe = tmp_path_factory.mktemp("njmodel") / "nj_bounded_scan_model"
    subprocess.run(["gcc", "-O2", "-o", str(exe), os.path.join(ROOT, "tools", "nj_bounded_scan_model.c"), "-lm"],
                   check=True, cwd=ROOT)
    return str(exe)


def _run(model, tmp_path, dist, *args):
    path = tmp_path / "d.f64"
    np.ascontiguousarray(dist, dtype=np.float64).tofile(path)
    p = subprocess.run([model, str(dist.shape[0]), str(path), *map(str, args)], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "mismatches 0;" in p.stdout
    return p.stdout


def _sym(a):
    d = (a + a.T) / 2
    np.fill_diagonal(d, 0.0)
    return d


@pytest.mark.parametrize("n,G,ctas,epoch", [(5, 8, 1, 1), (33, 8, 4, 2), (200, 16, 16, 8), (400, 64, 16, 8), (257, 32, 3, 5)])
def test_model_random_distances(model, tmp_path, n, G, ctas, epoch):
    rng = np.random.default_rng(n)
    _run(model, tmp_path, _sym(rng.random((n, n))), G, ctas, epoch)


def test_model_exact_ties_and_stars(model, tmp_path):
    """Rational distances with a common denominator (real Q ties) and star trees (everything ties): what cannot be
    excluded is evaluated exactly, so the lexicographic tie rule of the oracle decides."""
    rng = np.random.default_rng(99)
    for n in (120, 300):
        a = rng.integers(0, 6, size=(n, n)).astype(np.float64) / 8.0
        d = np.triu(a, 1)
        _run(model, tmp_path, d + d.T, 32, 16, 8)
        _run(model, tmp_path, d + d.T, 8, 2, 3)
    _run(model, tmp_path, np.ones((50, 50)) - np.eye(50), 32, 16, 8)
    _run(model, tmp_path, np.ones((130, 130)) - np.eye(130), 16, 4, 8)


def test_model_genotype_distances_and_relabelling(model, tmp_path):
    """Distances of the synthetic population sample (near-ties everywhere: the hard case for the bounds), with the
    samples relabelled at random too (tracked columns then collide in their groups)."""
    from oracle import cpu as oracle
    n, m = 150, 30000
    codes = oracle.synth_codes(n, 0, m, seed=7, miss_ppm=20000)
    counts = oracle.pair_counts(np.ascontiguousarray(codes.T), method="bits")
    for metric in ("p", "ibs"):
        dist, _ = oracle.normalise(counts, metric)
        out = _run(model, tmp_path, dist, 16, 16, 8)
        assert "rescans" in out
        _run(model, tmp_path, dist, 16, 16, 8, 11)


def test_model_tree_like_metric(model, tmp_path):
    """Additive (caterpillar) metric, symmetrised: the node merged last is joined again at once."""
    n = 120
    pos = np.cumsum(np.arange(1, n + 1, dtype=np.float64) * 0.01)
    height = np.linspace(0.5, 1.5, n)
    dist = np.abs(pos[:, None] - pos[None, :]) + height[:, None] + height[None, :]
    _run(model, tmp_path, _sym(dist), 32, 16, 8)
    _run(model, tmp_path, _sym(dist), 8, 16, 4, 5)
