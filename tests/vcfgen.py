"""Seeded VCF text generators for the parity tests (SNP and SV), covering the behavioural
edge cases SURVEY.md §4 lists for the reference converters. Only inputs on which the
unmodified reference neither raises nor desynchronises its rows are produced by the
``random_*`` functions; inputs that must be rejected live in ``bad_*``."""
from __future__ import annotations

import random

HEADER = ("##fileformat=VCFv4.2\n"
          "##source=phyloforge_b200-tests\n")


def _chrom_line(names):
    return "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(names) + "\n"


def sample_names(n, prefix="S"):
    return [f"{prefix}{i:04d}" for i in range(n)]


def _gt(rng, n_alleles, miss, phased_p=0.3, exotic=True):
    r = rng.random()
    if r < miss:
        if not exotic:
            return "./."
        return rng.choice(["./.", ".|.", ".", "./1", "0/.", "1", "0", "0/1/1", "x/y", "1|."])
    a = rng.randrange(n_alleles)
    b = rng.randrange(n_alleles)
    sep = "|" if rng.random() < phased_p else "/"
    return f"{a}{sep}{b}"


def random_snp_vcf(seed, n_samples, n_sites, miss=0.05, multi_p=0.1, star_p=0.03, indel_p=0.05,
                   lower_p=0.05, subfield_p=0.3, mono_p=0.02, crlf=False, blank_p=0.01):
    rng = random.Random(seed)
    names = sample_names(n_samples)
    out = [HEADER, _chrom_line(names)]
    eol = "\r\n" if crlf else "\n"
    for s in range(n_sites):
        ref = rng.choice("ACGT")
        others = [b for b in "ACGT" if b != ref]
        r = rng.random()
        if r < indel_p:
            alts = [ref + rng.choice("ACGT")] if rng.random() < 0.5 else ["<DEL>"]
            if rng.random() < 0.3:
                alts.append(rng.choice(others))
            if rng.random() < 0.3:
                ref = ref + "T"
        elif r < indel_p + multi_p:
            alts = rng.sample(others, rng.choice([2, 3]))
        elif r < indel_p + multi_p + star_p:
            alts = [rng.choice(others), "*"] if rng.random() < 0.5 else ["*"]
        elif r < indel_p + multi_p + star_p + mono_p:
            alts = ["."]
        else:
            alts = [rng.choice(others)]
        if rng.random() < lower_p:
            ref = ref.lower()
            alts = [a.lower() for a in alts]
        n_alleles = 1 + len(alts)
        with_sub = rng.random() < subfield_p
        fmt = "GT:DP:GQ" if with_sub else "GT"
        fields = []
        for _ in range(n_samples):
            g = _gt(rng, n_alleles, miss)
            if with_sub:
                g += f":{rng.randrange(1, 60)}:{rng.randrange(1, 99)}"
            fields.append(g)
        out.append(f"chr{1 + s % 5}\t{100 + s}\t.\t{ref}\t{','.join(alts)}\t.\tPASS\t.\t{fmt}\t"
                   + "\t".join(fields) + eol)
        if rng.random() < blank_p:
            out.append(eol)
    return "".join(out), names


def random_sv_vcf(seed, n_samples, n_sites, miss=0.05, hom_site_p=0.3, other_type_p=0.05,
                  subfield_p=0.3, no_svtype_p=0.02):
    rng = random.Random(seed)
    names = sample_names(n_samples, "V")
    out = [HEADER, _chrom_line(names)]
    for s in range(n_sites):
        r = rng.random()
        if r < other_type_p:
            typ = rng.choice(["DUP", "INV", "BND", "del"])
        else:
            typ = rng.choice(["DEL", "INS"])
        if rng.random() < no_svtype_p:
            info = f"END={1000 + s};SVLEN=55"
        elif rng.random() < 0.5:
            info = f"SVTYPE={typ};END={1000 + s};SVLEN={rng.randrange(50, 5000)}"
        else:
            info = f"END={1000 + s};SVTYPE={typ};IMPRECISE"   # flag AFTER SVTYPE is never split
        hom_site = rng.random() < hom_site_p
        with_sub = rng.random() < subfield_p
        fields = []
        for _ in range(n_samples):
            if rng.random() < miss:
                g = rng.choice(["./.", ".|."] if hom_site else ["./.", ".|.", "0/.", "./1", ".|0"])
            else:
                a = rng.randrange(2)
                b = a if hom_site else rng.randrange(2)
                g = f"{a}{'|' if rng.random() < 0.3 else '/'}{b}"
            if with_sub:
                g += f":{rng.randrange(1, 60)}"
            fields.append(g)
        out.append(f"chr{1 + s % 3}\t{500 + s}\tsv{s}\tN\t<{typ}>\t.\tPASS\t{info}\t"
                   f"{'GT:DP' if with_sub else 'GT'}\t" + "\t".join(fields) + "\n")
    return "".join(out), names


def edge_snp_vcf():
    """One line per row of the SURVEY.md §4 table that the reference survives."""
    names = ["a", "b", "c", "d", "e", "f"]
    rows = [
        # REF ALT  genotypes (6 samples)
        ("A", "G", ["0/0", "1/1", "0/1", "1|0", "./.", "."]),            # hom, het, missing
        ("A", "C", ["0/1", "0/0", "1/1", ".|1", "0", "0/1/1"]),          # half-missing, haploid, triploid
        ("A", "T", ["0/1", "1/0", "0|1", "1|1", "0/0", "./."]),
        ("C", "G", ["0/1", "1/1", "0/0", "0/1", "1/0", "0/0"]),
        ("C", "T", ["0/1", "1/1", "0/0", "0/1", "1/0", "0/0"]),
        ("G", "T", ["0/1", "1/1", "0/0", "0/1", "1/0", "0/0"]),
        ("a", "g", ["0/0", "1/1", "0/1", "0/0", "1/1", "0/1"]),          # lower-case alleles
        ("AT", "A", ["0/0", "1/1", "0/1", "0/0", "1/1", "0/1"]),         # indel: site skipped
        ("A", "<DEL>", ["0/0", "1/1", "0/1", "0/0", "1/1", "0/1"]),      # symbolic: skipped
        ("C", "G,T", ["1/2", "2/2", "0/2", "0/1", "1/1", "0/0"]),        # multi-allelic single-base
        ("A", "*", ["1/1", "0/1", "1/0", "0/0", "./.", "1|1"]),          # star allele
        ("G", "T,*", ["1/2", "2/2", "0/2", "0/1", "2/1", "0/0"]),
        ("T", ".", ["0/0", "0/1", "1/1", "1/0", "0/0", "./."]),          # monomorphic record
        ("A", "G", ["0/1:4", "1/1:12:99", "0/0:3,4", "./.:.", ".:.", "0|0:7"]),  # sub-fields
        ("A", "C,G,T", ["0/3", "3/3", "2/3", "1/2", "0/0", "3|1"]),
    ]
    out = [HEADER, "##contig=<ID=chr1>\n", _chrom_line(names)]
    for k, (ref, alt, gts) in enumerate(rows):
        out.append(f"chr1\t{10 + k}\t.\t{ref}\t{alt}\t50\tPASS\tDP=10\tGT\t" + "\t".join(gts) + "\n")
    out.append("\n")                                           # blank line is ignored
    out.append("   \n")
    out.append("chr1 999 . G A 50 PASS DP=1 GT 0/0 0/1 1/1 0/0 0/1 1/1\n")   # space-separated
    return "".join(out), names


def edge_sv_vcf():
    names = ["p", "q", "r", "s", "t"]
    rows = [
        ("SVTYPE=DEL;END=5", ["0/0", "1/1", "0/0", "1/1", "0/0"]),     # hom site, DEL map [1,0]
        ("SVTYPE=INS;END=5", ["0/0", "1/1", "0/0", "1/1", "0/0"]),     # hom site, INS map [0,1]
        ("SVTYPE=DEL", ["0/1", "1/1", "0/0", "1/0", "0|1"]),           # het site -> two columns
        ("SVTYPE=INS", ["0/1", "1/1", "0/0", "1/0", "1|0"]),
        ("SVTYPE=DUP;END=9", ["0/1", "1/1", "0/0", "1/0", "0/0"]),     # filtered type
        ("END=9;SVLEN=4", ["0/1", "1/1", "0/0", "1/0", "0/0"]),        # no SVTYPE key
        ("SVTYPE=DEL", ["./.", "1/1", "0/0", ".|.", "0/0"]),           # missing on a hom site
        ("SVTYPE=INS", ["./.", "0/1", "0|.", "./1", "1/1"]),           # missing on a het site
        ("END=3;SVTYPE=INS;IMPRECISE", ["0/0:3", "1/1:9", "0/0:1", "1/1:2", "1/1:8"]),
        ("SVTYPE=DEL", ["0/00", "1/1", "0/0", "1/1", "0/0"]),          # '0' != '00' -> het site
    ]
    out = [HEADER, _chrom_line(names)]
    for k, (info, gts) in enumerate(rows):
        out.append(f"chr2\t{100 + k}\tsv{k}\tN\t<SV>\t.\tPASS\t{info}\tGT\t" + "\t".join(gts) + "\n")
    return "".join(out), names


def bad_snp_lines():
    """(description, body line) pairs the product must reject (reference raises / corrupts)."""
    return [
        ("allele index out of range", "chr1\t5\t.\tA\tG\t.\t.\t.\tGT\t0/2\t0/0\n"),
        ("homozygous N allele", "chr1\t5\t.\tN\tG\t.\t.\t.\tGT\t0/0\t0/1\n"),
        ("empty GT subfield", "chr1\t5\t.\tA\tG\t.\t.\t.\tGT:DP\t:5\t0/1:3\n"),
        ("too few sample columns", "chr1\t5\t.\tA\tG\t.\t.\t.\tGT\t0/1\n"),
        ("too many sample columns", "chr1\t5\t.\tA\tG\t.\t.\t.\tGT\t0/1\t0/0\t1/1\n"),
    ]


def bad_sv_lines():
    return [
        ("flag before SVTYPE", "chr1\t5\t.\tN\t<DEL>\t.\t.\tIMPRECISE;SVTYPE=DEL\tGT\t0/1\t0/0\n"),
        ("allele index 2", "chr1\t5\t.\tN\t<DEL>\t.\t.\tSVTYPE=DEL\tGT\t2/2\t0/0\n"),
        ("haploid GT", "chr1\t5\t.\tN\t<DEL>\t.\t.\tSVTYPE=DEL\tGT\t1\t0/0\n"),
    ]


def write_bgzf(path, data: bytes, block=0xff00, seed=0):
    """bgzip-style file: independent gzip members with the 'BC' extra subfield (SAM specification 4.1), blocks of
    varying uncompressed size, terminated by the 28-byte empty EOF block."""
    import random
    import struct
    import zlib
    rng = random.Random(seed)

    def member(chunk: bytes) -> bytes:
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        cdata = co.compress(chunk) + co.flush()
        bsize = 12 + 6 + len(cdata) + 8
        assert bsize <= 65536
        return (b"\x1f\x8b\x08\x04" + b"\0\0\0\0" + b"\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1)
                + cdata + struct.pack("<II", zlib.crc32(chunk), len(chunk)))

    with open(path, "wb") as f:
        pos = 0
        while pos < len(data):
            k = rng.randint(1, block)
            f.write(member(data[pos:pos + k]))
            pos += k
        f.write(member(b""))
