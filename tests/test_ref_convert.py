"""The restated converters (oracle/ref_convert.py) against the UNMODIFIED reference: the
committed golden vectors always; fresh random VCFs too when /root/reference is present."""
import json
import os
import subprocess
import sys

import pytest

import vcfgen
from oracle import ref_convert as rc, run_reference

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SNP_CASES = ["snp_edge", "snp_rand_a", "snp_rand_b", "snp_rand_crlf", "snp_biallelic"]
SV_CASES = ["sv_edge", "sv_rand_a", "sv_rand_b"]


def test_manifest_matches_files(golden_dir):
    import hashlib
    man = json.load(open(os.path.join(golden_dir, "MANIFEST.json")))
    assert set(man) == set(SNP_CASES + SV_CASES)
    for name, rec in man.items():
        assert hashlib.sha256(open(os.path.join(golden_dir, f"{name}.vcf"), "rb").read()).hexdigest() == rec["input_sha256"]
        for fn, digest in rec["outputs"].items():
            assert hashlib.sha256(open(os.path.join(golden_dir, fn), "rb").read()).hexdigest() == digest


@pytest.mark.parametrize("case", SNP_CASES)
def test_snp_restatement_matches_golden(golden_dir, case):
    text = open(os.path.join(golden_dir, f"{case}.vcf"), "rb").read().decode()
    names, seqs, kept = rc.snp_convert(text)
    assert rc.fasta_text(names, seqs) == open(os.path.join(golden_dir, f"{case}.all_snp.fa")).read()
    assert rc.phylip_text(names, seqs) == open(os.path.join(golden_dir, f"{case}.all_snp.phy")).read()


@pytest.mark.parametrize("case", SV_CASES)
def test_sv_restatement_matches_golden(golden_dir, case):
    names, seqs = rc.sv_convert(open(os.path.join(golden_dir, f"{case}.vcf")).read())
    assert rc.fasta_text(names, seqs) == open(os.path.join(golden_dir, f"{case}.SV.matrix")).read()


def test_edge_table_of_survey_section_4(golden_dir):
    """Spot checks of the behaviours listed in SURVEY.md §4, read off the reference's golden output."""
    fa = open(os.path.join(golden_dir, "snp_edge.all_snp.fa")).read().split("\n")
    seq = dict(zip((h[1:] for h in fa[0::2]), fa[1::2]))
    assert seq["a"][0] == "A" and seq["b"][0] == "G" and seq["c"][0] == "R" and seq["d"][0] == "R"
    assert seq["e"][0] == "-" and seq["f"][0] == "-"                       # ./. and .
    assert seq["d"][1] == "-" and seq["e"][1] == "-" and seq["f"][1] == "-"  # .|1, haploid, triploid
    assert seq["a"][6] == "A" and seq["c"][6] == "R"                       # lower-case alleles upper-cased
    assert len(seq["a"]) == 14                                             # indel and <DEL> sites skipped
    assert seq["a"][7] == "K" and seq["b"][7] == "T"                       # C>G,T: 1/2 -> K, 2/2 -> T
    assert seq["a"][8] == "N" and seq["b"][8] == "a"                       # '*': hom -> N, het -> lower case
    assert seq["a"][10] == "T" and seq["b"][10] == "N" and seq["c"][10] == "-"   # ALT '.'
    mat = open(os.path.join(golden_dir, "sv_edge.SV.matrix")).read().split("\n")
    sv = dict(zip((h[1:] for h in mat[0::2]), mat[1::2]))
    assert sv["p"][:2] == "10" and sv["q"][:2] == "01"                     # DEL 0/0 -> 1, INS 0/0 -> 0
    assert sv["p"][2:4] == "10" and sv["s"][2:4] == "01"                   # het site: two columns, GT order


def test_iupac_table_of_baseline_md():
    rows = {"A": "AMRWaN", "C": "MCSYcN", "G": "RSGKgN", "T": "WYKTtN", "*": "acgtNN", ".": "NNNNN-"}
    for b1, row in rows.items():
        for b2, want in zip("ACGT*.", row):
            assert rc.iupac(b1, b2) == want, (b1, b2)
    with pytest.raises(rc.ReferenceWouldFail):
        rc.iupac("N", "N")


@pytest.mark.parametrize("desc,line", vcfgen.bad_snp_lines())
def test_bad_snp_lines_raise(desc, line):
    text = vcfgen.HEADER + vcfgen._chrom_line(["x", "y"]) + line
    with pytest.raises(rc.ReferenceWouldFail):
        rc.snp_convert(text)


@pytest.mark.parametrize("desc,line", vcfgen.bad_sv_lines())
def test_bad_sv_lines_raise(desc, line):
    text = vcfgen.HEADER + vcfgen._chrom_line(["x", "y"]) + line
    with pytest.raises(rc.ReferenceWouldFail):
        rc.sv_convert(text)


def _run_ref(*args):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "run_reference.py"), *map(str, args)],
                       capture_output=True, text=True)
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    return json.loads(lines[-1]) if lines else {"error": p.stderr[-500:]}


@pytest.mark.skipif(not run_reference.available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("seed", [11, 12])
def test_snp_restatement_vs_live_reference(tmp_path, seed):
    text, _ = vcfgen.random_snp_vcf(seed, 23, 250, miss=0.15, subfield_p=0.5)
    vcf = tmp_path / "in.vcf"
    vcf.write_text(text)
    res = _run_ref("snp", vcf, tmp_path / "out", 3, "treebest")
    names, seqs, _ = rc.snp_convert(text)
    assert open(res["out"]).read() == rc.fasta_text(names, seqs)


@pytest.mark.skipif(not run_reference.available(), reason="/root/reference not present (GPU box)")
def test_sv_restatement_vs_live_reference(tmp_path):
    text, _ = vcfgen.random_sv_vcf(13, 31, 200, miss=0.1)
    vcf = tmp_path / "in.vcf"
    vcf.write_text(text)
    res = _run_ref("sv", vcf, tmp_path / "out")
    names, seqs = rc.sv_convert(text)
    assert open(res["out"]).read() == rc.fasta_text(names, seqs)


@pytest.mark.skipif(not run_reference.available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("kind,lines", [("snp", vcfgen.bad_snp_lines()), ("sv", vcfgen.bad_sv_lines())])
def test_live_reference_fails_or_corrupts_on_bad_lines(tmp_path, kind, lines):
    """The inputs we reject really are ones the unmodified reference cannot convert: it raises,
    or (empty GT / wrong column count) silently writes rows of unequal length."""
    for k, (desc, line) in enumerate(lines):
        good = ("chr1\t1\t.\tA\tG\t.\t.\t.\tGT\t0/1\t1/1\n" if kind == "snp"
                else "chr1\t1\t.\tN\t<DEL>\t.\t.\tSVTYPE=DEL\tGT\t0/1\t1/1\n")
        vcf = tmp_path / f"bad{k}.vcf"
        vcf.write_text(vcfgen.HEADER + vcfgen._chrom_line(["x", "y"]) + good + line + good)
        res = _run_ref(kind, vcf, tmp_path / f"out{k}", *((2, "treebest") if kind == "snp" else ()))
        if "error" in res:
            continue
        seqs = open(res["out"]).read().split("\n")[1::2]
        assert len(set(len(s) for s in seqs if s is not None)) > 1 or "haploid" in desc, desc
