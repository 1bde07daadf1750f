"""Bounded-scan NJ kernel (csrc/nj_lazy.cu) against the full-scan kernel and the CPU oracle: same join list and branch
lengths bit for bit, rescans / list evaluations per iteration, time.  python tools/nj_lazy_check.py [n:sites ...]"""
import ctypes as C
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import cpu as oracle
from phyloforge_b200._lib import check
from phyloforge_b200.engine import Engine

eng = Engine(0)


def genotype_dist(n, m, miss=0, metric="p"):
    counts = eng.alloc_counts(n)
    chunk = 1 << 20
    nw = chunk // 32
    planes = eng.alloc_planes(n, nw)
    for s0 in range(0, m, chunk):
        k = min(chunk, m - s0)
        codes = eng.synth_codes(n, s0, k, seed=42, miss_ppm=miss)
        eng.pack_codes(codes, n, planes, nw, n_sites=k)
        eng.pair_counts(planes, n, nw, counts, word_end=(k + 31) // 32)
    dist, _ = eng.normalise(counts, metric)
    return dist


def timed_nj(d, mode, reps=3):
    check(eng.lib.pf_nj_profile(mode, None))
    m, l = eng.nj(d)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        wd = d.clone()
        e0.record()
        m, l = eng.nj(wd, destroy=True)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out = (C.c_ulonglong * 8)()
    check(eng.lib.pf_nj_profile(0, out))
    return m, l, best, list(out)


def run(name, d, with_oracle=True):
    n = d.shape[0]
    m1, l1, t1, st1 = timed_nj(d, 8)
    m0, l0, t0, _ = timed_nj(d, 4)
    same = bool(torch.equal(m0, m1) and torch.equal(l0, l1))
    calls = 4
    rec = {"case": name, "n": n, "lazy_ms": round(t1, 3), "full_scan_ms": round(t0, 3), "same_as_full_scan": same,
           "joined_by_lazy": st1[7] // calls if st1[7] else 0,
           "rescans_per_iter": round(st1[4] / max(1, st1[6]), 3), "list_evals_per_iter": round(st1[5] / max(1, st1[6]), 3),
           "us_per_iter": round(t1 * 1e3 / n, 3)}
    # phase accounting (instrumented kernel)
    check(eng.lib.pf_nj_lazy_profile(1, None))
    eng.nj(d)
    torch.cuda.synchronize()
    ph = (C.c_ulonglong * 16)()
    check(eng.lib.pf_nj_lazy_profile(0, ph))
    names = ["-", "A champions+T", "B checks+list evals", "C rescans+publish", "D cluster barrier", "D winner", "E update", "F rows", "barrier alone"]
    rec["phases_us_per_iter"] = {nm: round(ph[i] / 1e3 / n, 3) for i, nm in enumerate(names)}
    if ph[12]:
        rec["rank0_rescan_us"] = {"n": int(ph[12]), "loads+scan": round(ph[9] / 1e3 / ph[12], 3),
                                  "reduce+barrier": round(ph[10] / 1e3 / ph[12], 3), "merge+list+barrier": round(ph[11] / 1e3 / ph[12], 3), "one_load_round_trip": round(ph[13] / 1e3 / ph[12], 3)}
    if with_oracle and n <= 3000:
        wm, wl = oracle.nj(d.cpu().numpy())
        rec["same_as_oracle"] = bool(np.array_equal(wm, m1.cpu().numpy()) and np.array_equal(wl, l1.cpu().numpy()))
    print(json.dumps(rec), flush=True)
    return rec


def run_batch(k, n):
    """k matrices of n nodes joined by one pf_nj_batch call: bounded scan (one cluster per matrix) vs full scan."""
    a = torch.rand((k, n, n), dtype=torch.float64, device=eng.device)
    d = (a + a.transpose(1, 2)) / 2
    d.diagonal(dim1=1, dim2=2).zero_()
    res = {}
    for mode in (8, 4):
        check(eng.lib.pf_nj_profile(mode, None))
        m, l = eng.nj_batch(d.clone())
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            dd = d.clone()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            m, l = eng.nj_batch(dd)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        res[mode] = (m, l, best)
    check(eng.lib.pf_nj_profile(0, None))
    print(json.dumps({"case": f"batch {k} x {n}", "lazy_ms": round(res[8][2], 3), "full_scan_ms": round(res[4][2], 3),
                      "ms_per_tree_lazy": round(res[8][2] / k, 3), "ms_per_tree_full_scan": round(res[4][2] / k, 3),
                      "same": bool(torch.equal(res[8][0], res[4][0]) and torch.equal(res[8][1], res[4][1]))}), flush=True)


cases = sys.argv[1:] or ["u:64", "u:700", "u:1500", "g:1000:1000000", "g:3000:1000000", "u:3000"]
for c in cases:
    parts = c.split(":")
    if parts[0] == "b":
        run_batch(int(parts[1]), int(parts[2]))
        continue
    if parts[0] == "u":
        n = int(parts[1])
        a = torch.rand((n, n), dtype=torch.float64, device=eng.device)
        d = (a + a.T) / 2
        d.fill_diagonal_(0)
        run(c, d)
    else:
        n, m = int(parts[1]), int(parts[2])
        miss = int(parts[3]) if len(parts) > 3 else 0
        metric = parts[4] if len(parts) > 4 else "p"
        run(c, genotype_dist(n, m, miss, metric))
