"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's VCF -> character-matrix
converters, written from the behaviour of treetool/vcftree.py and treetool/svtree.py (each
function cites the lines it follows). Pinned against the UNMODIFIED reference by
tests/golden/ (and, where /root/reference exists, by fresh random inputs in
tests/test_ref_convert.py). See oracle/__init__.py for who may import this.

Where the reference raises (and the run ends with "Failed to Convert ...", exit 1) this
module raises ReferenceWouldFail with the same cause; where the reference silently
desynchronises its rows (a bug) it raises too — the product rejects the same inputs.
"""
from __future__ import annotations

import numpy as np

_BASES = "ACGT"
_HET = {frozenset("AC"): "M", frozenset("AG"): "R", frozenset("AT"): "W",
        frozenset("CG"): "S", frozenset("CT"): "Y", frozenset("GT"): "K"}


class ReferenceWouldFail(Exception):
    """Input on which the unmodified reference raises or corrupts its output."""


def iupac(b1: str, b2: str) -> str:
    """VcfTree.generate_ambiguous_code (vcftree.py:26-65) as a decision table.
    '.' is the converter's stand-in for an unparsable genotype (vcftree.py:128-129)."""
    b1, b2 = b1.upper(), b2.upper()                     # vcftree.py:31-32
    if b1 == b2:                                        # vcftree.py:35-41
        if b1 in _BASES and len(b1) == 1:
            return b1
        if b1 == ".":
            return "-"
        if b1 == "*":
            return "N"
        raise ReferenceWouldFail(f"homozygous allele {b1!r}: reference returns None")
    in1, in2 = b1 in _BASES and len(b1) == 1, b2 in _BASES and len(b2) == 1
    if not in1 and not in2:                             # vcftree.py:42-43
        return "N"
    if in1 and in2:                                     # vcftree.py:44-55
        return _HET[frozenset((b1, b2))]
    base, other = (b1, b2) if in1 else (b2, b1)
    if other == "*":                                    # vcftree.py:56-63
        return base.lower()
    return "N"                                          # vcftree.py:64-65


def _parse_snp_gt(gt: str, alleles):
    """vcftree.py:124-129: two integers separated by '/' or '|', anything else -> ('.','.').
    An index past the allele list is an IndexError the reference does not catch."""
    parts = gt.replace("|", "/").split("/")
    if len(parts) != 2:
        return ".", "."
    try:
        a1, a2 = int(parts[0]), int(parts[1])
    except ValueError:
        return ".", "."
    if a1 < 0 or a2 < 0 or not parts[0].isascii() or not parts[1].isascii() \
            or not parts[0].isdigit() or not parts[1].isdigit():
        raise ReferenceWouldFail(f"exotic integer in GT {gt!r} (signed / non-ASCII digits)")
    if a1 >= len(alleles) or a2 >= len(alleles):
        raise ReferenceWouldFail(f"allele index out of range in GT {gt!r}")
    return alleles[a1], alleles[a2]


def snp_convert(text: str):
    """VcfTree.convert_vcf_to_seq + merge_seq on the whole file (vcftree.py:96-171):
    returns (sample names, list of per-sample strings, list of kept (REF, ALT-string))."""
    names, seqs, kept = None, None, []
    for line in text.split("\n"):
        if line.startswith("#CHROM"):                   # vcftree.py:99-102
            names = line.strip().split()[9:]
            seqs = [[] for _ in names]
            continue
        if line.startswith("#") or not line.strip():    # vcftree.py:104-106
            continue
        if names is None:
            raise ReferenceWouldFail("record before #CHROM header")
        f = line.strip().split()                        # vcftree.py:107
        if len(f) < 10:
            raise ReferenceWouldFail("fewer than 10 columns")
        alleles = [f[3]] + f[4].split(",")              # vcftree.py:108-109
        if any(len(a) > 1 for a in alleles):            # vcftree.py:113-119
            continue
        if len(f) - 9 != len(names):
            raise ReferenceWouldFail("sample column count differs from header")
        col = []
        for field in f[9:]:                             # vcftree.py:120-133
            gt = field.split(":")[0]
            if not gt:
                raise ReferenceWouldFail("empty GT subfield: reference desynchronises rows")
            b1, b2 = _parse_snp_gt(gt, alleles)
            col.append(iupac(b1, b2))
        for i, ch in enumerate(col):
            seqs[i].append(ch)
        kept.append((f[3], f[4]))
    if names is None:
        raise ReferenceWouldFail("no #CHROM header")
    return names, ["".join(s) for s in seqs], kept


def fasta_text(names, seqs) -> str:
    """vcftree.py:168-171 / svtree.py:122-125."""
    return "".join(f">{n}\n{s}\n" for n, s in zip(names, seqs))


def phylip_text(names, seqs) -> str:
    """vcftree.py:159-166."""
    head = f"{len(names)} {len(seqs[0]) if seqs else 0}\n"
    return head + "".join(f"{n}  {s}\n" for n, s in zip(names, seqs))


def snp_fits_2bit(ref: str, alt: str) -> bool:
    """A kept site is representable by codes {0,1,2,3} iff biallelic single-base A/C/G/T
    (ALT may also be '.', a monomorphic record) with REF != ALT."""
    r, a = ref.upper(), alt.upper()
    return len(r) == 1 and r in _BASES and "," not in alt and len(a) == 1 and (a in _BASES or a == ".") and r != a


def snp_codes(seqs, kept):
    """Characters -> 2-bit codes, sample-major u8 [n, m_fit] over the sites that fit 2 bits.
    char == REF -> 0, == ALT -> 2, IUPAC het -> 1, '-' / 'N' -> 3."""
    cols = [k for k, (r, a) in enumerate(kept) if snp_fits_2bit(r, a)]
    out = np.full((len(seqs), len(cols)), 3, dtype=np.uint8)
    for c, k in enumerate(cols):
        ref, alt = kept[k][0].upper(), kept[k][1].upper()
        for i, s in enumerate(seqs):
            ch = s[k]
            if ch == ref:
                out[i, c] = 0
            elif ch == alt:
                out[i, c] = 2
            elif ch in "MRWSYK":
                out[i, c] = 1
            elif ch in "-N":
                out[i, c] = 3
            else:
                raise AssertionError(f"unexpected character {ch!r} at a biallelic site")
    return out, cols


# ------------------------------------------------------------------ SV
def _sv_split(gt: str):
    """svtree.py:72-77: '/' takes precedence over '|'; pieces 0 and 1 of the split."""
    if "/" in gt:
        p = gt.split("/")
    elif "|" in gt:
        p = gt.split("|")
    else:
        raise ReferenceWouldFail(f"GT {gt!r} without separator: reference re-uses stale alleles")
    return p[0], p[1]


def _sv_char(allele: str, amap) -> str:
    """svtree.py:96-98 / 112-118: int(allele) indexes the DEL/INS map; ValueError -> '.'."""
    try:
        v = int(allele)
    except ValueError:
        return "."
    if not (allele.isascii() and allele.isdigit()):
        raise ReferenceWouldFail(f"exotic integer {allele!r} in SV GT")
    if v >= 2:
        raise ReferenceWouldFail("SV allele index >= 2: IndexError in the reference")
    return str(amap[v])


def sv_convert(text: str):
    """SvTree.convert_vcf_to_seq (svtree.py:14-125): returns (names, per-sample strings)."""
    names, seqs = None, None
    for line in text.split("\n"):
        if line.startswith("#CHROM"):                   # svtree.py:38-41
            names = line.strip().split()[9:]
            seqs = [[] for _ in names]
            continue
        if line.startswith("#") or not line.strip():
            continue
        if names is None:
            raise ReferenceWouldFail("record before #CHROM header")
        f = line.strip().split()
        if len(f) < 10:
            raise ReferenceWouldFail("fewer than 10 columns")
        sv_type = None
        for pair in f[7].split(";"):                    # svtree.py:49-55
            kv = pair.split("=")
            if len(kv) != 2:
                raise ReferenceWouldFail(f"INFO item {pair!r} before SVTYPE has no single '='")
            if kv[0] == "SVTYPE":
                sv_type = kv[1]
                break
        if sv_type not in ("DEL", "INS"):               # svtree.py:56-57
            continue
        amap = (1, 0) if sv_type == "DEL" else (0, 1)   # svtree.py:58-61
        if len(f) - 9 != len(names):
            raise ReferenceWouldFail("sample column count differs from header")
        pairs = []
        for field in f[9:]:
            gt = field.split(":")[0]
            if not gt:
                raise ReferenceWouldFail("empty GT subfield")
            pairs.append(_sv_split(gt))
        hom = all(a1 == a2 for a1, a2 in pairs)         # svtree.py:66-82 (string compare)
        for i, (a1, a2) in enumerate(pairs):
            if hom:                                     # svtree.py:84-99
                seqs[i].append(_sv_char(a1, amap))
            else:                                       # svtree.py:100-120
                seqs[i].append(_sv_char(a1, amap) + _sv_char(a2, amap))
    if names is None:
        raise ReferenceWouldFail("no #CHROM header")
    return names, ["".join(s) for s in seqs]


def sv_codes(seqs):
    """'0' -> 0, '1' -> 2, '.' -> 3, sample-major u8 [n, m]."""
    lut = np.full(256, 255, dtype=np.uint8)
    lut[ord("0")], lut[ord("1")], lut[ord(".")] = 0, 2, 3
    if not seqs:
        return np.zeros((0, 0), dtype=np.uint8)
    arr = np.frombuffer("".join(seqs).encode(), dtype=np.uint8).reshape(len(seqs), -1)
    out = lut[arr]
    assert (out != 255).all()
    return out
