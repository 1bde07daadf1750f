"""Where the time of `phyloforge -s` goes on VCF text (1,000 samples x M SNPs by default; C3 = 1M): cold first run and
best of the following ones, the conversion alone (vcf.load_snp_vcf with / without the character matrix), the
alignment-file writer with 1 / 4 / 8 threads, the distance-matrix writer, and the whole entry point.
usage: r02_entry_profile.py [RECORDS] [SAMPLES]"""
import json, os, sys, time, tempfile, shutil
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from phyloforge_b200 import phylip, run, synth, vcf
from phyloforge_b200.engine import Engine

m = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
eng = Engine(0)
names = [f"S{i:04d}" for i in range(n)]
codes = eng.synth_codes(n, 0, m, seed=42)[:, :n].cpu().numpy()
text = synth.snp_vcf_bytes(codes, names)[0]
del codes
d = tempfile.mkdtemp(dir="/dev/shm")
out = {"samples": n, "records": m, "vcf_bytes": len(text), "host_threads": os.cpu_count()}


def best(fn, reps=3):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return ts, r


try:
    src = os.path.join(d, "in.vcf")
    open(src, "wb").write(text)
    gb = len(text) / 1e9
    ts, packed = best(lambda: vcf.load_snp_vcf(eng, src, keep_chars=True))
    out["load_with_chars_s"] = ts
    out["load_with_chars_GBps"] = gb / min(ts)
    ts, _ = best(lambda: vcf.load_snp_vcf(eng, src, keep_chars=False))
    out["load_planes_only_s"] = ts
    out["load_planes_only_GBps"] = gb / min(ts)
    for th in (1, 4, 8):
        ts, _ = best(lambda: vcf.write_fasta(os.path.join(d, "x.fa"), packed.names, packed.chars, threads=th))
        out[f"write_fasta_threads{th}_s"] = ts
    out["fasta_bytes"] = os.path.getsize(os.path.join(d, "x.fa"))
    dist = np.random.default_rng(0).random((n, n))
    ts, _ = best(lambda: phylip.write_distance_matrix(os.path.join(d, "x.dist"), names, dist))
    out["write_dist_native_s"] = ts
    if n <= 1500:
        ts, _ = best(lambda: phylip.write_distance_matrix_py(os.path.join(d, "y.dist"), names, dist), reps=1)
        out["write_dist_python_s"] = ts
    del packed
    cfg = os.path.join(d, "run.cfg")
    open(cfg, "w").write(f"[snp_opt]\nvcf_file = {src}\nout_path = {d}/out\ntree_software = treebest\nthread = 8\n")
    ts, _ = best(lambda: run.main(["-s", cfg]))
    rep = json.load(open(f"{d}/out/SNP.phyloforge_b200.json"))
    out["entry_point_s"] = ts
    out["entry_point_GBps"] = gb / min(ts)
    out["entry_point_report"] = {k: rep[k] for k in ("convert_seconds", "tree_seconds")}
finally:
    shutil.rmtree(d, ignore_errors=True)
print(json.dumps(out, indent=1), file=sys.__stdout__)
