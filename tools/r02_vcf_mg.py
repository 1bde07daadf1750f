"""C3-shaped VCF TEXT (1,000 samples x M SNPs; M = 1M is BASELINE configs[2], ~4 GB of text) through `phyloforge -s`
with gpus = 1 and gpus = N, for both worker models (threads of one process; processes + NCCL process group) and,
for threads, both ways of summing the counts (in-process NCCL; peer copies): wall time of the whole entry point
(workers start, convert their byte range, count, allreduce, rank 0 writes), best of `reps` runs in one process (the
first run of every variant also pays the CUDA contexts of the devices it is the first to touch), the conversion and
tree seconds the run itself reports, and whether every output file equals the one-GPU run's.

usage: r02_vcf_mg.py WORLD [RECORDS] [variants, comma separated: threads-nccl,threads-p2p,processes] [reps] [BOOTSTRAP]
BOOTSTRAP > 0 adds `bootstrap = BOOTSTRAP` (seeded block-bootstrap support, the replicates dealt out over the workers)."""
import json, os, sys, time, tempfile, shutil
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phyloforge_b200 import multigpu, run, synth
from phyloforge_b200.engine import Engine

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n, m = 1000, int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
variants = (sys.argv[3] if len(sys.argv) > 3 else "threads-nccl,threads-p2p,processes").split(",")
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
boot = int(sys.argv[5]) if len(sys.argv) > 5 else 0
eng = Engine(0)
names = [f"S{i:04d}" for i in range(n)]
codes = eng.synth_codes(n, 0, m, seed=42)[:, :n].cpu().numpy()
text = synth.snp_vcf_bytes(codes, names)[0]
del codes
d = tempfile.mkdtemp(dir="/dev/shm")
out = {"samples": n, "records": m, "vcf_bytes": len(text), "world": world, "reps": reps, "bootstrap": boot}
FILES = ("all_snp.fa", "SNP.dist", "SNP_treebest.NHX")


def one(tag, extra, env=None):
    cfg = os.path.join(d, f"run_{tag}.cfg")
    open(cfg, "w").write(f"[snp_opt]\nvcf_file = {src}\nout_path = {d}/out_{tag}\ntree_software = treebest\nthread = 8\n{extra}"
                         + (f"bootstrap = {boot}\nbootstrap_seed = 7\n" if boot else ""))
    old = {k: os.environ.get(k) for k in (env or {})}
    os.environ.update(env or {})
    try:
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            run.main(["-s", cfg])
            ts.append(time.perf_counter() - t0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    rep = json.load(open(f"{d}/out_{tag}/SNP.phyloforge_b200.json"))
    files = [open(f"{d}/out_{tag}/{f}", "rb").read() for f in FILES]
    shutil.rmtree(f"{d}/out_{tag}", ignore_errors=True)
    return {"wall_s": min(ts), "wall_s_runs": ts, "convert_s": rep["convert_seconds"], "tree_s": rep["tree_seconds"],
            "text_GBps": len(text) / min(ts) / 1e9}, files


try:
    src = os.path.join(d, "in.vcf")
    open(src, "wb").write(text)
    out["gpus_1"], ref = one("1", "gpus = 1\n")
    same = {}
    for v in variants:
        if v == "processes":
            rec, files = one(v, f"gpus = {world}\ngpu_workers = processes\n")
        else:
            how = v.split("-")[1]
            rec, files = one(v, f"gpus = {world}\ngpu_workers = threads\n", {multigpu.ENV_THREAD_REDUCE: how})
        out[f"gpus_{world}_{v}"] = rec
        same[v] = files == ref
    out["identical_outputs"] = same
finally:
    shutil.rmtree(d, ignore_errors=True)
print(json.dumps(out, indent=1), file=sys.__stdout__)
