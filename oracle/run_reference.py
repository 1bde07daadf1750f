"""TEST INFRASTRUCTURE ONLY — run the UNMODIFIED reference converters from /root/reference.

The reference's own half of the hot path (VCF -> character matrix) is pure Python and runs
here once ``Bio`` is replaced by oracle/bio_shim (SURVEY.md §8(c) recipe). This script is
used (a) by tests/golden/make_golden.py to generate the committed golden vectors, (b) by
the CPU tests, when /root/reference exists, to check oracle/ref_convert.py against the real
thing on fresh random inputs, and (c) by ``bench.py --impl reference`` when the reference
tree is present. /root/reference does not exist on the GPU box: nothing run there may
depend on this file succeeding.

It must run in its own process: the reference parses ``sys.argv`` at import time
(treetool/script.py:31) and keeps module-level state (treetool/vcftree.py:12).

    python oracle/run_reference.py snp <vcf> <out_dir> <thread> <treebest|iqtree|...>
    python oracle/run_reference.py sv  <vcf> <out_dir>
prints one JSON line with timings and output paths.
"""
import json
import os
import sys
import time
import warnings

REFERENCE_ROOT = os.environ.get("PHYLOFORGE_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "treetool", "vcftree.py"))


def _write_cfg(path, section, vcf, out_dir, thread, tool):
    with open(path, "w") as f:
        f.write(f"[{section}]\nvcf_file = {vcf}\nout_path = {out_dir}\n"
                f"tree_software = {tool}\nthread = {thread}\n")


def run_snp(vcf, out_dir, thread, tool):
    import multiprocessing
    os.makedirs(out_dir, exist_ok=True)
    cfg = os.path.join(out_dir, "ref_run.cfg")
    _write_cfg(cfg, "snp_opt", vcf, out_dir, thread, tool)
    sys.argv = ["phyloforge", "-s", cfg]
    sys.path[:0] = [os.path.join(HERE, "bio_shim"), REFERENCE_ROOT]
    warnings.simplefilter("ignore")
    import treetool.vcftree as vt  # the reference, unmodified

    t = vt.VcfTree()
    t0 = time.perf_counter()
    chunks = t.cutFile(t.vcf_file, int(t.thread), t.working_dir)          # vcftree.py:216
    t1 = time.perf_counter()
    with multiprocessing.Pool(int(t.thread)) as pool:                       # vcftree.py:219-220
        pool.map(t.convert_vcf_to_seq, chunks)
    t2 = time.perf_counter()
    fa_list = t.change_file_extension(chunks, ".fa")                        # vcftree.py:217
    if t.tree_software.upper() in ("TREEBEST", "FASTTREE"):                 # vcftree.py:229-235
        t.merge_seq(fa_list, f"{t.out_path}/all_snp", "fa")
        out = f"{t.out_path}/all_snp.fa"
    else:
        t.merge_seq(fa_list, f"{t.out_path}/all_snp", "phylip")
        out = f"{t.out_path}/all_snp.phy"
    t3 = time.perf_counter()
    return {"out": out, "cut_s": t1 - t0, "convert_s": t2 - t1, "merge_s": t3 - t2,
            "total_s": t3 - t0, "thread": int(t.thread)}


def run_sv(vcf, out_dir):
    os.makedirs(out_dir, exist_ok=True)
    cfg = os.path.join(out_dir, "ref_run.cfg")
    _write_cfg(cfg, "sv_opt", vcf, out_dir, 2, "iqtree")
    sys.argv = ["phyloforge", "-S", cfg]
    sys.path[:0] = [os.path.join(HERE, "bio_shim"), REFERENCE_ROOT]
    warnings.simplefilter("ignore")
    import treetool.svtree as st  # the reference, unmodified

    t = st.SvTree()
    out = f"{t.out_path}/SV.matrix"
    t0 = time.perf_counter()
    t.convert_vcf_to_seq(t.vcf_file, out)                                   # svtree.py:131
    t1 = time.perf_counter()
    return {"out": out, "convert_s": t1 - t0, "total_s": t1 - t0, "thread": 1}


def main(argv):
    if not available():
        print(json.dumps({"unavailable": f"{REFERENCE_ROOT} not present"}))
        return 3
    kind = argv[1]
    try:
        if kind == "snp":
            res = run_snp(argv[2], argv[3], int(argv[4]), argv[5])
        elif kind == "sv":
            res = run_sv(argv[2], argv[3])
        else:
            raise SystemExit(f"unknown kind {kind}")
    except SystemExit:
        raise
    except BaseException as e:  # noqa: BLE001 - report how the reference failed
        print(json.dumps({"error": type(e).__name__, "message": str(e)}))
        return 1
    print(json.dumps(res))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
