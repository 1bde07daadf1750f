"""VCF ingestion: file bytes -> GPU parser -> bit planes (+ the reference's character matrix).

Host side of the converters the reference runs in Python:
  VcfTree.cutFile / convert_vcf_to_seq / merge_seq     treetool/vcftree.py:67-171
  SvTree.convert_vcf_to_seq                            treetool/svtree.py:14-125
The host only finds the ``#CHROM`` header and streams the file to the device in chunks cut
at line boundaries; line splitting, field splitting, genotype decoding, site filtering and
packing all run in libphyloforge_b200.so (csrc/vcf_parse.cu, csrc/pack.cu).
"""
from __future__ import annotations

import ctypes as C
import mmap
import os
from dataclasses import dataclass

import numpy as np
import torch

from ._lib import PhyloForgeError, check

ERR_TEXT = {
    1: "wrong number of columns (fewer than 10 fields, or sample count differs from the #CHROM header)",
    2: "allele index out of range (IndexError in the reference)",
    3: "homozygous genotype of an allele that is not A/C/G/T, '.' or '*' (reference fails)",
    4: "empty GT sub-field (the reference silently desynchronises its rows)",
    5: "more than 8 single-base alleles",
    6: "INFO item without exactly one '=' before SVTYPE (ValueError in the reference)",
    7: "GT without '/' or '|' (the reference re-uses the previous sample's alleles)",
    8: "signed / non-ASCII / underscore integer in GT",
}

DEFAULT_CHUNK_BYTES = 256 << 20


class VcfFormatError(PhyloForgeError):
    """The input is one the unmodified reference fails on; message names line and cause."""


@dataclass
class PackedGenotypes:
    names: list
    n_samples: int
    n_sites: int            # matrix columns packed into `planes`
    n_words: int            # words per tile in `planes` (>= ceil(n_sites / 32))
    planes: torch.Tensor    # int32, tiled bit planes (include/phyloforge_b200.h)
    chars: np.ndarray | None  # uint8 [n_samples, n_columns]: the reference's character matrix
    n_columns: int          # columns of the reference's matrix (== n_sites unless sites were left out)
    n_lines: int
    n_not_2bit: int = 0     # kept SNP sites that do not fit the 2-bit code (multi-allelic, '*')
    h2d_bytes: int = 0


def read_header(path: str):
    """Sample names from the first line that starts with ``#CHROM`` (vcftree.py:69-73,
    99-100). Returns (names, file_size)."""
    size = os.path.getsize(path)
    names = None
    with open(path, "rb") as f:
        for raw in f:
            if raw.startswith(b"#CHROM"):
                names = raw.decode("utf-8", "replace").strip().split()[9:]
                break
            if not raw.startswith(b"#") and raw.strip():
                break
    if names is None:
        raise VcfFormatError(f"{path}: no #CHROM header line before the first record")
    if not names:
        raise VcfFormatError(f"{path}: the #CHROM header names no samples")
    return names, size


def _chunk_bounds(mm, size: int, chunk_bytes: int):
    """Cut [0, size) into chunks that end just after a newline."""
    bounds, start = [], 0
    while start < size:
        end = min(size, start + chunk_bytes)
        if end < size:
            nl = mm.rfind(b"\n", start, end)
            if nl < 0:
                nl = mm.find(b"\n", end)
                end = size if nl < 0 else nl + 1
            else:
                end = nl + 1
        bounds.append((start, end))
        start = end
    return bounds


def _load(engine, path: str, kind: str, keep_chars: bool, non_biallelic: str, chunk_bytes: int):
    lib, dev = engine.lib, engine.device
    names, size = read_header(path)
    n = len(names)
    rpl = 1 if kind == "snp" else 2          # matrix rows a line can yield
    ld = (n + 3) // 4 * 4
    stream = torch.cuda.current_stream().cuda_stream
    if size == 0:
        raise VcfFormatError(f"{path}: empty file")

    f = open(path, "rb")
    mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
    view = np.frombuffer(mm, dtype=np.uint8)
    try:
        total_lines_ub = int(np.count_nonzero(view == 10)) + 1
        n_words_cap = max(1, (total_lines_ub * rpl + 31) // 32)
        planes = engine.alloc_planes(n, n_words_cap)
        bounds = _chunk_bounds(mm, size, chunk_bytes)
        max_chunk = max(e - s for s, e in bounds)
        max_lines = 0
        for s, e in bounds:
            max_lines = max(max_lines, int(np.count_nonzero(view[s:e] == 10)) + 1)

        d_text = torch.empty(max_chunk, dtype=torch.uint8, device=dev)
        h_pin = torch.empty(max_chunk, dtype=torch.uint8).pin_memory()
        d_lines = torch.empty(max_lines + 1, dtype=torch.int64, device=dev)
        d_info = torch.empty(max_lines, dtype=torch.int32, device=dev)
        d_work = torch.empty(int(lib.pf_line_index_work_bytes(max_chunk)), dtype=torch.uint8, device=dev)
        # rows [0, 32) of the code buffer carry the sites left over from the previous chunk
        carry_rows = 32
        d_codes = torch.empty((carry_rows + max_lines * rpl, ld), dtype=torch.uint8, device=dev)
        d_chars = torch.empty((max_lines * rpl, ld), dtype=torch.uint8, device=dev)
        d_rows = torch.empty(carry_rows + max_lines * rpl, dtype=torch.int64, device=dev)
        d_rows_all = torch.empty(max_lines * rpl, dtype=torch.int64, device=dev) if keep_chars else None

        n_carry = 0            # leftover sites (< 32) waiting in d_codes[0:n_carry]
        words_done = 0
        n_lines_total = 0
        n_unfit_total = 0
        h2d = 0
        char_blocks = []
        line_base = 0
        summary = (C.c_int64 * 5)()
        parse = lib.pf_vcf_snp_parse if kind == "snp" else lib.pf_vcf_sv_parse
        codes_body = d_codes[carry_rows:]

        for (s, e) in bounds:
            nb = e - s
            h_pin[:nb].numpy()[:] = view[s:e]
            d_text[:nb].copy_(h_pin[:nb], non_blocking=True)
            h2d += nb
            n_lines = C.c_int64()
            check(lib.pf_line_index(d_text.data_ptr(), nb, d_lines.data_ptr(), max_lines + 1, C.byref(n_lines),
                                    d_work.data_ptr(), stream))
            nl = n_lines.value
            check(parse(d_text.data_ptr(), nb, d_lines.data_ptr(), nl, n, d_chars.data_ptr(),
                        codes_body.data_ptr(), ld, d_info.data_ptr(), stream))
            require_fit = 1 if kind == "snp" else 0
            # gather list for packing: leftover rows first, then this chunk's surviving rows
            check(lib.pf_vcf_select_rows(d_info.data_ptr(), nl, rpl, require_fit,
                                         d_rows[carry_rows:].data_ptr(), max_lines * rpl, summary, stream))
            n_cols, n_err, n_unfit, first_line, first_cls = (int(x) for x in summary)
            if n_err:
                raise VcfFormatError(
                    f"{path}: {n_err} record(s) the reference cannot convert; first at body line "
                    f"{line_base + first_line + 1}: {ERR_TEXT.get(first_cls, first_cls)}")
            if n_unfit and non_biallelic != "skip":
                raise VcfFormatError(
                    f"{path}: {n_unfit} kept site(s) are not biallelic A/C/G/T and do not fit the 2-bit "
                    "genotype code (set `non_biallelic = skip` in the config to leave them out of the "
                    "distance matrix; all_snp.fa still contains them)")
            n_unfit_total += n_unfit
            if keep_chars:
                if n_unfit:
                    check(lib.pf_vcf_select_rows(d_info.data_ptr(), nl, rpl, 0, d_rows_all.data_ptr(),
                                                 max_lines * rpl, summary, stream))
                    n_all = int(summary[0])
                    rows_all = d_rows_all[:n_all]
                else:
                    n_all = n_cols
                    rows_all = d_rows[carry_rows:carry_rows + n_cols]
                if n_all:
                    block = engine.transpose_u8(d_chars, n, rows=rows_all, n_rows=n_all)
                    char_blocks.append(block.cpu().numpy())
            # pack: full words now, remainder carried to the next chunk
            d_rows[carry_rows:carry_rows + n_cols] += carry_rows        # rows index into d_codes
            first = carry_rows - n_carry
            if n_carry:
                d_rows[first:carry_rows] = torch.arange(0, n_carry, dtype=torch.int64, device=dev)
            avail = n_carry + n_cols
            full = (avail // 32) * 32
            if full:
                engine.pack_codes(d_codes, n, planes, n_words_cap, word0=words_done,
                                  rows=d_rows[first:first + full], n_sites=full)
                words_done += full // 32
            rem = avail - full
            if rem:
                idx = d_rows[first + full:first + avail]
                d_codes[:rem] = d_codes.index_select(0, idx)
            n_carry = rem
            n_lines_total += nl
            line_base += nl
        if n_carry:
            rows = torch.arange(0, n_carry, dtype=torch.int64, device=dev)
            engine.pack_codes(d_codes, n, planes, n_words_cap, word0=words_done, rows=rows, n_sites=n_carry)
        n_sites = words_done * 32 + n_carry
        torch.cuda.current_stream().synchronize()
    finally:
        del view                 # the mmap cannot close while a numpy view of it is alive
        mm.close()
        f.close()

    chars = None
    n_columns = n_sites
    if keep_chars:
        chars = (np.concatenate(char_blocks, axis=1) if char_blocks
                 else np.zeros((n, 0), dtype=np.uint8))
        n_columns = chars.shape[1]
    return PackedGenotypes(names=names, n_samples=n, n_sites=n_sites, n_words=n_words_cap, planes=planes,
                           chars=chars, n_columns=n_columns, n_lines=n_lines_total,
                           n_not_2bit=n_unfit_total, h2d_bytes=h2d)


def load_snp_vcf(engine, path: str, keep_chars: bool = True, non_biallelic: str = "error",
                 chunk_bytes: int = DEFAULT_CHUNK_BYTES) -> PackedGenotypes:
    """SNP VCF -> planes of the biallelic sites + the IUPAC character matrix of every kept site."""
    return _load(engine, path, "snp", keep_chars, non_biallelic, chunk_bytes)


def load_sv_vcf(engine, path: str, keep_chars: bool = True,
                chunk_bytes: int = DEFAULT_CHUNK_BYTES) -> PackedGenotypes:
    """SV VCF -> planes and the 0/1/. matrix (one or two columns per DEL/INS record)."""
    return _load(engine, path, "sv", keep_chars, "error", chunk_bytes)


def write_fasta(path: str, names, chars: np.ndarray):
    """``>id\\nseq\\n`` per sample (vcftree.py:168-171, svtree.py:122-125)."""
    with open(path, "wb") as f:
        for name, row in zip(names, chars):
            f.write(b">" + name.encode() + b"\n")
            f.write(row.tobytes())
            f.write(b"\n")


def write_phylip(path: str, names, chars: np.ndarray):
    """``N L`` then ``id  seq`` per sample (vcftree.py:159-166)."""
    with open(path, "wb") as f:
        f.write(f"{len(names)} {chars.shape[1]}\n".encode())
        for name, row in zip(names, chars):
            f.write(name.encode() + b"  ")
            f.write(row.tobytes())
            f.write(b"\n")
