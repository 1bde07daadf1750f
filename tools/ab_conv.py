"""A/B of the planes -> E2M1 operand-stream conversion (PRMT expansion vs shared-memory LUT) on 3,000 x 10M:
conversion / GEMM milliseconds from the CUDA events inside pf_pair_counts_tc, and bit-identical counts."""
import ctypes as C, json, statistics, sys
import torch
sys.path.insert(0, ".")
from phyloforge_b200._lib import check
from phyloforge_b200.engine import Engine

eng = Engine(0)
n, m = 3000, int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
codes = eng.synth_codes(n, 0, m, seed=42)
nw = eng.n_words(m)
planes = eng.alloc_planes(n, nw)
eng.pack_codes(codes, n, planes, nw)
del codes
out, ref = {}, None
for name, bits in (("x4_prmt", 0), ("x1_lut", 2048 << 4), ("x1_prmt", 4096 << 4), ("x4_prmt_again", 0)):
    check(eng.lib.pf_pair_counts_tc_profile(bits, None))
    check(eng.lib.pf_pair_counts_tc_timing(1, None))
    conv, gemm = [], []
    for it in range(5):
        counts = eng.alloc_counts(n)
        eng.pair_counts(planes, n, nw, counts, variant=3)
        ms2 = (C.c_float * 2)()
        check(eng.lib.pf_pair_counts_tc_timing(1, ms2))
        if it:
            conv.append(float(ms2[0])); gemm.append(float(ms2[1]))
    check(eng.lib.pf_pair_counts_tc_timing(0, None))
    if ref is None:
        ref = counts.clone()
    out[name] = {"conversion_ms": statistics.mean(conv), "gemm_ms": statistics.mean(gemm), "same_counts": bool(torch.equal(ref, counts))}
check(eng.lib.pf_pair_counts_tc_profile(0, None))
bytes_moved = planes.numel() * 4 * 3   # planes read + 2x their size written as nibbles
for k, v in out.items():
    v["conversion_gbps"] = bytes_moved / (v["conversion_ms"] * 1e-3) / 1e9
print(json.dumps(out, indent=1))
