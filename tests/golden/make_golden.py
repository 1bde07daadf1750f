"""Generate the committed golden vectors by running the UNMODIFIED reference converters
(/root/reference, through oracle/run_reference.py and the Biopython stand-in) on seeded
VCFs from tests/vcfgen.py. Run here (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

For each case NAME it writes NAME.vcf (input) and the reference's output:
NAME.all_snp.fa / NAME.all_snp.phy (treetool/vcftree.py:145-171) or NAME.SV.matrix
(treetool/svtree.py:122-125), plus MANIFEST.json with sha256 sums and the thread counts
used (the SNP output is thread-count invariant; two counts are run and compared).
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
sys.path.insert(0, str(ROOT / "tests"))
import vcfgen  # noqa: E402

RUNNER = ROOT / "oracle" / "run_reference.py"


def run_ref(*args):
    p = subprocess.run([sys.executable, str(RUNNER), *map(str, args)], capture_output=True, text=True)
    line = [ln for ln in p.stdout.strip().splitlines() if ln.startswith("{")]
    if p.returncode != 0 or not line:
        raise RuntimeError(f"reference run failed rc={p.returncode}\n{p.stdout}\n{p.stderr[-2000:]}")
    return json.loads(line[-1])


def sha(path):
    return hashlib.sha256(Path(path).read_bytes()).hexdigest()


def main():
    cases_snp = {
        "snp_edge": vcfgen.edge_snp_vcf()[0],
        "snp_rand_a": vcfgen.random_snp_vcf(101, 12, 300)[0],
        "snp_rand_b": vcfgen.random_snp_vcf(202, 67, 150, miss=0.2, subfield_p=0.6)[0],
        "snp_rand_crlf": vcfgen.random_snp_vcf(303, 5, 80, crlf=True)[0],
        "snp_biallelic": vcfgen.random_snp_vcf(404, 33, 257, miss=0.1, multi_p=0, star_p=0, indel_p=0,
                                               lower_p=0, mono_p=0)[0],
    }
    cases_sv = {
        "sv_edge": vcfgen.edge_sv_vcf()[0],
        "sv_rand_a": vcfgen.random_sv_vcf(111, 9, 200)[0],
        "sv_rand_b": vcfgen.random_sv_vcf(222, 70, 120, miss=0.15)[0],
    }
    manifest = {}
    for name, text in cases_snp.items():
        vcf = HERE / f"{name}.vcf"
        vcf.write_text(text, newline="")
        outs = {}
        for thread, tool in ((2, "treebest"), (5, "treebest"), (3, "iqtree")):
            with tempfile.TemporaryDirectory() as td:
                res = run_ref("snp", vcf, td, thread, tool)
                ext = ".fa" if tool == "treebest" else ".phy"
                dst = HERE / f"{name}.all_snp{ext}"
                data = Path(res["out"]).read_bytes()
                if dst.name in outs:
                    assert outs[dst.name] == data, f"{name}: output differs between thread counts"
                outs[dst.name] = data
                dst.write_bytes(data)
        manifest[name] = {"kind": "snp", "input_sha256": sha(vcf),
                          "outputs": {k: hashlib.sha256(v).hexdigest() for k, v in outs.items()},
                          "threads_checked": [2, 5, 3]}
    for name, text in cases_sv.items():
        vcf = HERE / f"{name}.vcf"
        vcf.write_text(text, newline="")
        with tempfile.TemporaryDirectory() as td:
            res = run_ref("sv", vcf, td)
            dst = HERE / f"{name}.SV.matrix"
            shutil.copyfile(res["out"], dst)
        manifest[name] = {"kind": "sv", "input_sha256": sha(vcf), "outputs": {dst.name: sha(dst)}}
    (HERE / "MANIFEST.json").write_text(json.dumps(manifest, indent=1, sort_keys=True) + "\n")
    print(f"wrote {len(manifest)} golden cases to {HERE}")


if __name__ == "__main__":
    main()
