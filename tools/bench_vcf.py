"""Throughput of the GPU VCF converters (rows a3-a6, a8 of SURVEY.md §8(a)) on synthetic VCF text.
    python tools/bench_vcf.py > gpurun_out/bench_vcf.json
The reference's own numbers for the same step: SNP 50 x 100k in 2.97 s on 8 cores
(1.68 M genotypes/s); SV 500 x 100k in 766.5 s on 1 core (65 k genotypes/s)."""
import json, os, sys, tempfile, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cpu as oracle
from phyloforge_b200 import synth, vcf
from phyloforge_b200.engine import Engine

eng = Engine(0)
out = []
for kind, n, m, miss in (("snp", 50, 100_000, 0), ("snp", 1000, 200_000, 0), ("sv", 500, 100_000, 50_000)):
    names = [f"S{i:04d}" for i in range(n)]
    codes = oracle.synth_codes(n, 0, m, seed=42, miss_ppm=miss)
    text = (synth.snp_vcf_bytes(codes, names)[0] if kind == "snp" else synth.sv_vcf_bytes(codes, names)[0])
    with tempfile.NamedTemporaryFile(suffix=".vcf", delete=False, dir="/dev/shm") as f:
        f.write(text)
        path = f.name
    load = vcf.load_snp_vcf if kind == "snp" else vcf.load_sv_vcf
    rec = {"kind": kind, "samples": n, "records": m, "text_bytes": len(text)}
    for keep in (True, False):
        load(eng, path, keep_chars=keep)                      # warm-up
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            p = load(eng, path, keep_chars=keep)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        t = min(ts)
        rec["with_chars" if keep else "planes_only"] = {"seconds": t, "text_GBps": len(text) / t / 1e9,
                                                         "genotypes_per_s": n * m / t}
    os.unlink(path)
    out.append(rec)
print(json.dumps(out, indent=1))
