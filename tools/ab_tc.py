"""A/B of the tensor-core pair-count kernels: impl 1 (single CTA, 128 x 192 tiles) against impl 2 (CTA pair,
cta_group::2, 256 x 240 tiles) and its pipeline geometries. Parity against the oracle on small shapes,
then CUDA-event timings on a large one. Diagnostic."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import cpu as oracle
from phyloforge_b200._lib import check
from phyloforge_b200.engine import Engine

eng = Engine(0)
cfgs = [(1, 0, 0), (2, 0, 0), (2, 1, 0), (2, 2, 0), (2, 3, 0), (2, 5, 0)]     # (impl, cfg, debug bits)
if len(sys.argv) > 3:
    cfgs = [tuple(int(x) for x in a.split(":")) for a in sys.argv[3].split(",")]
skip_parity = len(sys.argv) > 4 and sys.argv[4] == "noparity"
rng = np.random.default_rng(7)
bad = 0
for n, m, kind in () if skip_parity else ((5, 70, 0), (120, 900, 0), (121, 4000, 2), (240, 5000, 0), (241, 3000, 1), (256, 7000, 2), (257, 2100, 0),
                   (500, 6000, 0), (513, 9000, 2), (777, 3001, 3), (1000, 20000, 0), (1300, 12000, 2)):
    codes = rng.integers(0, 3, size=(m, n), dtype=np.uint8)
    if kind == 1:
        codes[rng.random(m) < 0.05] = 3
    elif kind == 2:
        codes[rng.random((m, n)) < 0.1] = 3
    elif kind == 3:
        codes[:, 3] = 3
    d = torch.from_numpy(codes).to(eng.device)
    nw = eng.n_words(m)
    planes = eng.alloc_planes(n, nw)
    eng.pack_codes(d, n, planes, nw)
    want = oracle.pair_counts(np.ascontiguousarray(codes.T), method="bits")
    cut = int(rng.integers(0, nw + 1))
    for impl, cfg, dbg in cfgs:
        check(eng.lib.pf_pair_counts_tc_set_impl(impl, cfg))
        check(eng.lib.pf_pair_counts_tc_profile(dbg << 4, None))
        c = eng.alloc_counts(n)
        eng.pair_counts(planes, n, nw, c, word_begin=0, word_end=cut, variant=3)
        eng.pair_counts(planes, n, nw, c, word_begin=cut, word_end=nw, variant=3)
        torch.cuda.synchronize()
        got = c.cpu().numpy().astype(np.int64)
        ok = np.array_equal(got, want)
        if not ok:
            bad += 1
            w = np.argwhere(got != want)
            print(f"MISMATCH n={n} m={m} kind={kind} impl={impl} cfg={cfg} dbg={dbg}: {len(w)} entries, first {w[:4].tolist()}", flush=True)
    print(f"parity n={n} m={m} kind={kind}: done", flush=True)
print("parity:", "OK" if bad == 0 else f"{bad} FAILURES", flush=True)

n, m = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3000, 10_000_000)
codes = eng.synth_codes(n, 0, m, seed=42)
nw = eng.n_words(m)
planes = eng.alloc_planes(n, nw)
eng.pack_codes(codes, n, planes, nw)
del codes
torch.cuda.empty_cache()
ref = None
for impl, cfg, dbg in cfgs:
    check(eng.lib.pf_pair_counts_tc_set_impl(impl, cfg))
    check(eng.lib.pf_pair_counts_tc_profile(dbg << 4, None))
    counts = eng.alloc_counts(n)
    eng.pair_counts(planes, n, nw, counts, variant=3)
    torch.cuda.synchronize()
    if ref is None:
        ref = counts.clone()
    same = bool(torch.equal(ref, counts))
    ts = []
    for _ in range(4):
        counts.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.pair_counts(planes, n, nw, counts, variant=3)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    buf = (C.c_longlong * 16)()
    eng.lib.pf_pair_counts_tc_profile(1 | (dbg << 4), None)
    check(eng.lib.pf_pair_counts_tc_timing(1, None))
    eng.pair_counts(planes, n, nw, counts, variant=3)
    torch.cuda.synchronize()
    ms2 = (C.c_float * 2)()
    check(eng.lib.pf_pair_counts_tc_timing(0, ms2))
    eng.lib.pf_pair_counts_tc_profile(0, buf)
    v = list(buf)
    steps = max(v[8], 1)
    print(f"impl={impl} cfg={cfg} dbg={dbg} n={n} m={m}: counts (conversion + GEMM) {min(ts):.2f} ms (all {[round(t, 2) for t in ts]}), same as first={same}; "
          f"per step: MMA thread {v[1] / steps:.0f} clk, waiting {v[0] / steps:.0f}; A warp {v[4] / steps:.0f}, waiting {v[3] / steps:.0f}; "
          f"B warp {v[11] / steps:.0f}, waiting {v[10] / steps:.0f} (expansion warps of CTA rank {(dbg >> 6) & 1}); "
          f"proxy fence A {v[12] / steps:.0f} B {v[14] / steps:.0f}, end-of-stage barrier A {v[13] / steps:.0f} B {v[15] / steps:.0f}; "
          f"instrumented run: conversion {ms2[0]:.2f} ms, GEMM {ms2[1]:.2f} ms, MMA-thread loop time summed over items / pairs = "
          f"{v[1] / 74 / 1.965e6:.2f} ms", flush=True)
