"""C3-shaped VCF TEXT (1,000 samples x 200k SNPs, ~0.8 GB) through `phyloforge -s` with gpus = 1 and gpus = N:
wall time of the whole entry point (workers start, parse their byte range, count, allreduce, rank 0 writes)."""
import json, os, sys, time, tempfile, shutil
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from phyloforge_b200 import run, synth
from phyloforge_b200.engine import Engine

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n, m = 1000, int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
eng = Engine(0)
names = [f"S{i:04d}" for i in range(n)]
codes = eng.synth_codes(n, 0, m, seed=42)[:, :n].cpu().numpy()
text = synth.snp_vcf_bytes(codes, names)[0]
d = tempfile.mkdtemp(dir="/dev/shm")
out = {"samples": n, "records": m, "vcf_bytes": len(text)}
try:
    src = os.path.join(d, "in.vcf")
    open(src, "wb").write(text)
    files = {}
    for g in (1, world):
        cfg = os.path.join(d, f"run{g}.cfg")
        open(cfg, "w").write(f"[snp_opt]\nvcf_file = {src}\nout_path = {d}/out{g}\ntree_software = treebest\nthread = 8\ngpus = {g}\n")
        ts = []
        for _ in range(2):
            t0 = time.perf_counter()
            run.main(["-s", cfg])
            ts.append(time.perf_counter() - t0)
        rep = json.load(open(f"{d}/out{g}/SNP.phyloforge_b200.json"))
        out[f"gpus_{g}"] = {"wall_s": min(ts), "convert_s": rep["convert_seconds"], "tree_s": rep["tree_seconds"]}
        files[g] = [open(f"{d}/out{g}/{f}", "rb").read() for f in ("all_snp.fa", "SNP.dist", "SNP_treebest.NHX")]
    out["identical_outputs"] = files[1] == files[world]
finally:
    shutil.rmtree(d, ignore_errors=True)
print(json.dumps(out, indent=1), file=sys.__stdout__)
