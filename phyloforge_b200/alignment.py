"""Character-matrix ingestion: the FASTA / PHYLIP files the reference's converters write
(``all_snp.fa`` / ``all_snp.phy``: treetool/vcftree.py:159-171; ``SV.matrix``:
treetool/svtree.py:122-125) back into packed genotypes, so that the tree step can be driven from
the same file `treebest nj` receives (treetool/script.py:359-360). The file is read on the host
(it is sample-major text); transposition, character -> code mapping, column filtering and packing
run on the GPU."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._lib import check
from .vcf import PackedGenotypes, VcfFormatError


def read_alignment(path: str):
    """(names, uint8 matrix [n, m]) from FASTA (``>id`` + sequence lines) or the reference's PHYLIP
    flavour (``N L`` then ``id  seq``)."""
    with open(path, "rb") as f:
        data = f.read()
    if not data.strip():
        raise VcfFormatError(f"{path}: empty alignment")
    names, seqs = [], []
    if data.lstrip().startswith(b">"):
        cur = None
        for line in data.split(b"\n"):
            line = line.rstrip(b"\r")
            if line.startswith(b">"):
                parts = line[1:].split()
                names.append(parts[0].decode() if parts else "")
                cur = []
                seqs.append(cur)
            elif cur is not None and line.strip():
                cur.append(line.strip())
        seqs = [b"".join(s) for s in seqs]
    else:
        lines = [ln for ln in data.split(b"\n") if ln.strip()]
        head = lines[0].split()
        if len(head) != 2 or not head[0].isdigit():
            raise VcfFormatError(f"{path}: neither FASTA nor PHYLIP")
        for ln in lines[1:1 + int(head[0])]:
            parts = ln.split()
            names.append(parts[0].decode())
            seqs.append(b"".join(parts[1:]))
    if not seqs:
        raise VcfFormatError(f"{path}: no sequences")
    m = len(seqs[0])
    if any(len(s) != m for s in seqs):
        raise VcfFormatError(f"{path}: sequences of unequal length (a desynchronised matrix)")
    mat = np.frombuffer(b"".join(seqs), dtype=np.uint8).reshape(len(seqs), m).copy()
    return names, mat


def load_alignment(engine, path: str, non_biallelic: str = "include") -> PackedGenotypes:
    names, mat = read_alignment(path)
    n, m = mat.shape
    dev = engine.device
    stream = torch.cuda.current_stream().cuda_stream
    lib = engine.lib
    d_mat = torch.from_numpy(np.ascontiguousarray(mat)).to(dev)                  # sample-major [n, m]
    d_cols = engine.transpose_u8(d_mat, m, n_rows=n)                             # site-major [m, n]
    ld = (n + 3) // 4 * 4
    d_codes = torch.empty((max(m, 1), ld), dtype=torch.uint8, device=dev)
    d_info = torch.empty(max(m, 1), dtype=torch.int32, device=dev)
    check(lib.pf_chars_to_codes(d_cols.data_ptr(), d_cols.stride(0), m, n, d_codes.data_ptr(), ld,
                                d_info.data_ptr(), stream))
    d_rows = torch.empty(max(m, 1), dtype=torch.int64, device=dev)
    summary = (C.c_int64 * 5)()
    check(lib.pf_vcf_select_rows(d_info.data_ptr(), m, 1, 1, d_rows.data_ptr(), m, summary, stream))
    n_fit, n_unfit = int(summary[0]), int(summary[2])
    if n_unfit and non_biallelic not in ("skip", "include"):
        raise VcfFormatError(
            f"{path}: {n_unfit} column(s) hold more than two alleles (or '*'-derived characters) and do not fit "
            "the 2-bit genotype code (non_biallelic = skip leaves them out of the distance matrix)")
    n_words = max(1, (n_fit + 31) // 32)
    planes = engine.alloc_planes(n, n_words)
    if n_fit:
        engine.pack_codes(d_codes, n, planes, n_words, rows=d_rows[:n_fit], n_sites=n_fit)
    else:
        planes.fill_(-1)
    extra = None
    if n_unfit and non_biallelic == "include":
        is_unfit = torch.ones(m, dtype=torch.bool, device=dev)
        is_unfit[d_rows[:n_fit]] = False
        extra = d_cols.index_select(0, is_unfit.nonzero().flatten())
    return PackedGenotypes(names=names, n_samples=n, n_sites=n_fit, n_words=n_words, planes=planes, chars=mat,
                           n_columns=m, n_lines=m, n_not_2bit=n_unfit, h2d_bytes=int(mat.nbytes), extra_chars=extra)
