"""Host-side driver of the hot path: device buffers (PyTorch tensors, used only as
memory + streams) and the calls into libphyloforge_b200.so.

Stages (reference lines they replace):
  pack       site-major genotype codes -> bit planes      vcftree.py:120-133, svtree.py:84-120
  counts     all-pairs V / D1 / D2 integer matrices       `treebest nj` (script.py:359-360)
  allreduce  sum of the site shards' count matrices       (new: sites sharded over GPUs)
  normalise  counts -> fp64 p-distance / IBS distance     `treebest nj`
  nj         neighbor-joining join order + branch lengths `treebest nj`

There is no CPU path: every stage raises if the CUDA library or a GPU is missing.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib
from ._lib import PhyloForgeError, check
from .newick import merges_to_newick, merges_to_newick_native, two_leaf_newick

TILE = 64
METRICS = {"p": 0, "pdist": 0, "p-distance": 0, 0: 0, "ibs": 1, "IBS": 1, 1: 1}


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


@dataclass
class TreeResult:
    newick: str
    merge: np.ndarray
    length: np.ndarray
    dist: torch.Tensor | None = None
    counts: torch.Tensor | None = None
    n_zero_valid: int = 0
    timings_ms: dict = field(default_factory=dict)
    d2h_bytes: int = 0
    dist_host: np.ndarray | None = None


class Engine:
    """One engine per process / per GPU (one process per GPU under torchrun)."""

    def __init__(self, device: int | None = None):
        if not torch.cuda.is_available():
            raise PhyloForgeError("no CUDA device: phyloforge_b200 has no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        torch.cuda.set_device(self.device)
        sm, maj, mnr = C.c_int(), C.c_int(), C.c_int()
        mem, l2 = C.c_int64(), C.c_int()
        check(self.lib.pf_device_info(C.byref(sm), C.byref(maj), C.byref(mnr), C.byref(mem), C.byref(l2)))
        self.sm_count, self.cc = sm.value, (maj.value, mnr.value)
        self.total_mem, self.l2_bytes = mem.value, l2.value
        if self.cc[0] != 10:
            raise PhyloForgeError(f"compute capability {self.cc} is not sm_100: this library is B200-only")

    # ------------------------------------------------------------------ buffers
    @staticmethod
    def n_words(n_sites: int) -> int:
        return (n_sites + 31) // 32

    def alloc_planes(self, n_samples: int, n_words: int) -> torch.Tensor:
        n = int(self.lib.pf_planes_len(n_samples, n_words))
        return torch.empty(n, dtype=torch.int32, device=self.device)

    def alloc_counts(self, n_samples: int) -> torch.Tensor:
        return torch.zeros((3, n_samples, n_samples), dtype=torch.int32, device=self.device)

    # ------------------------------------------------------------------ stages
    def synth_codes(self, n_samples: int, site0: int, n_sites: int, seed: int = 42,
                    n_pops: int | None = None, miss_ppm: int = 0, out: torch.Tensor | None = None):
        """Site-major u8 codes [n_sites, ld] generated on the device."""
        if n_pops is None:
            n_pops = max(2, n_samples // 25)
        ld = (n_samples + 3) // 4 * 4
        if out is None:
            out = torch.empty((n_sites, ld), dtype=torch.uint8, device=self.device)
        check(self.lib.pf_synth_codes(out.data_ptr(), n_samples, out.stride(0), site0, n_sites, seed, n_pops,
                                      miss_ppm, _stream_ptr()))
        return out

    def pack_codes(self, codes: torch.Tensor, n_samples: int, planes: torch.Tensor, n_words: int,
                   word0: int = 0, rows: torch.Tensor | None = None, n_sites: int | None = None):
        """codes: site-major u8 [rows, ld]; packs n_sites sites starting at word0."""
        assert codes.dtype == torch.uint8 and codes.is_cuda and codes.stride(1) == 1
        if n_sites is None:
            n_sites = int(rows.numel()) if rows is not None else codes.shape[0]
        if rows is not None:
            assert rows.dtype == torch.int64 and rows.is_cuda
        check(self.lib.pf_pack_codes_u8(codes.data_ptr(), codes.stride(0), _lib.ptr(rows), n_samples, n_sites,
                                        planes.data_ptr(), n_words, word0, _stream_ptr()))

    def pack_codes_2bit(self, codes2: torch.Tensor, n_samples: int, planes: torch.Tensor, n_words: int,
                        word0: int = 0, n_sites: int | None = None, plink: bool = False):
        """codes2: site-major u8 [rows, ld2] holding 4 samples per byte (2 bits each)."""
        assert codes2.dtype == torch.uint8 and codes2.is_cuda and codes2.stride(1) == 1
        if n_sites is None:
            n_sites = codes2.shape[0]
        check(self.lib.pf_pack_codes_2bit(codes2.data_ptr(), codes2.stride(0), n_samples, n_sites, planes.data_ptr(),
                                          n_words, word0, 1 if plink else 0, _stream_ptr()))

    @staticmethod
    def to_2bit_rows(codes: torch.Tensor, n_samples: int) -> torch.Tensor:
        """Site-major u8 codes [m, >= n] -> site-major 2-bit rows [m, ld2] (ld2 % 4 == 0; padding = code 3).
        Input-format helper for tests and benchmarks (torch ops, not a timed path)."""
        m = codes.shape[0]
        n16 = (n_samples + 15) // 16 * 16
        c = torch.full((m, n16), 3, dtype=torch.uint8, device=codes.device)
        c[:, :n_samples] = codes[:, :n_samples] & 3
        c = c.view(m, n16 // 4, 4)
        return (c[..., 0] | (c[..., 1] << 2) | (c[..., 2] << 4) | (c[..., 3] << 6)).contiguous()

    def unpack_codes(self, planes: torch.Tensor, n_samples: int, n_words: int, n_sites: int,
                     word0: int = 0) -> torch.Tensor:
        """Sample-major u8 codes [n_samples, n_sites]."""
        out = torch.empty((n_samples, max(n_sites, 1)), dtype=torch.uint8, device=self.device)
        check(self.lib.pf_unpack_codes_u8(planes.data_ptr(), n_words, word0, n_samples, n_sites, out.data_ptr(),
                                          out.stride(0), _stream_ptr()))
        return out[:, :n_sites]

    def transpose_u8(self, mat: torch.Tensor, n_cols: int, rows: torch.Tensor | None = None,
                     n_rows: int | None = None) -> torch.Tensor:
        """out[i, k] = mat[rows[k], i]: site-major characters -> per-sample sequences."""
        if n_rows is None:
            n_rows = int(rows.numel()) if rows is not None else mat.shape[0]
        out = torch.empty((n_cols, max(n_rows, 1)), dtype=torch.uint8, device=self.device)
        check(self.lib.pf_transpose_u8(mat.data_ptr(), mat.stride(0), _lib.ptr(rows), n_rows, n_cols,
                                       out.data_ptr(), out.stride(0), _stream_ptr()))
        return out[:, :n_rows]

    # pair-count kernel variants (include/phyloforge_b200.h): 1 plain popcount, 2 carry-save popcount,
    # 3 tcgen05 int8 indicator GEMMs (one launch; a second one when genotypes are missing in some
    # samples only); 0 = auto: 3 when the sample count fills a 128 x 192 tensor-core tile reasonably,
    # else 2.
    TC_MIN_SAMPLES = 192
    # the tensor-core variant's scratch (two 4-bit operand streams over padded tiles) is ~4x the planes of
    # the word range it is given: longer ranges are walked in pieces of at most this many bytes of scratch,
    # accumulating into the same count matrices (integer adds: the result does not depend on the cut)
    TC_WORK_BUDGET = 56 << 30

    def pair_counts(self, planes: torch.Tensor, n_samples: int, n_words: int, counts: torch.Tensor,
                    word_begin: int = 0, word_end: int | None = None, variant: int = 0,
                    tc_general: bool = True) -> int:
        """counts[3, n, n] int32 += V / D1 / D2 over words [word_begin, word_end). Returns the
        variant that ran. tc_general=False makes variant 3 decline data with sample-specific
        missingness (the popcount kernel then runs, and the call waits for the device once);
        self.tc_launches = GEMM launches that did work in the last variant-3 call."""
        if word_end is None:
            word_end = n_words
        assert counts.dtype == torch.int32 and counts.is_contiguous() and counts.shape[0] == 3
        if variant == 0:
            variant = 3 if n_samples >= self.TC_MIN_SAMPLES else 2
        if variant == 3 and word_end - word_begin > 1024:
            budget = min(self.TC_WORK_BUDGET, max(1 << 30, self.total_mem // 3))
            need = int(self.lib.pf_pair_counts_tc_work_bytes(n_samples, word_end - word_begin))
            if need > budget:
                pieces = -(-need // budget)
                step = max(1024, -(-(word_end - word_begin) // pieces + 63) // 64 * 64)
                ran, general = 3, False
                for wb in range(word_begin, word_end, step):
                    ran = self.pair_counts(planes, n_samples, n_words, counts, wb, min(word_end, wb + step), variant,
                                           tc_general)
                    general = general or (ran == 3 and self.tc_launches == 2)
                self._tc_general_any = general
                return ran
        self._tc_general_any = None
        if variant == 3:
            need = int(self.lib.pf_pair_counts_tc_work_bytes(n_samples, word_end - word_begin))
            work = getattr(self, "_tc_work", None)
            if work is None or work.numel() < need:
                self._tc_work = None
                work = self._tc_work = torch.empty(need, dtype=torch.uint8, device=self.device)
            used = C.c_int(0)
            check(self.lib.pf_pair_counts_tc(planes.data_ptr(), n_samples, n_words, word_begin, word_end,
                                             counts.data_ptr(), counts.stride(1), work.data_ptr(), work.numel(),
                                             1 if tc_general else 0, C.byref(used), _stream_ptr()))
            self._tc_last = (n_samples, word_end - word_begin) if used.value else None
            if used.value:
                return 3
            variant = 2
        check(self.lib.pf_pair_counts(planes.data_ptr(), n_samples, n_words, word_begin, word_end,
                                      counts.data_ptr(), counts.stride(1), variant, _stream_ptr()))
        return variant

    def set_tc_impl(self, impl: int = 2, cfg: int = 0, strict: bool = False) -> None:
        """A/B switch of the tensor-core pair-count kernel (results are identical): impl 1 = one CTA per
        128 x 192 tile pair, impl 2 = a CTA pair (tcgen05.mma.cta_group::2) per 256-row tile pair (default);
        cfg = tile width / pipeline geometry of the pair kernel; strict = cluster-scope release on every
        operand hand-over (the slow, formally scoped form; see csrc/pair_counts_tc.cu)."""
        check(self.lib.pf_pair_counts_tc_set_impl(impl, cfg))
        check(self.lib.pf_pair_counts_tc_profile((1 << 4) if strict else 0, None))

    @property
    def tc_launches(self) -> int:
        """GEMM launches that did work in the last variant-3 call: 2 when it met genotypes missing in
        some samples only, else 1 (0: the call declined). Synchronises the current stream."""
        if getattr(self, "_tc_general_any", None) is not None:      # a call that was walked in pieces
            return 2 if self._tc_general_any else 1
        last = getattr(self, "_tc_last", None)
        if last is None or last[1] <= 0:
            return 0 if last is None else 1
        g = C.c_int(0)
        check(self.lib.pf_pair_counts_tc_was_general(self._tc_work.data_ptr(), last[0], last[1], C.byref(g),
                                                     _stream_ptr()))
        return 2 if g.value else 1

    # ------------------------------------------------------------------ bootstrap support
    def block_counts(self, planes: torch.Tensor, n_samples: int, n_words: int, n_blocks: int,
                     word_end: int | None = None, variant: int = 0, word0: int = 0, total_words: int | None = None,
                     n_extra_blocks: int = 0) -> torch.Tensor:
        """Count matrices of n_blocks contiguous word ranges: int32 [n_blocks + n_extra_blocks, 3, n, n] (the extra
        blocks are left zero for the caller). The ranges are cut on the GLOBAL word axis [0, total_words); this
        buffer holds global words [word0, word0 + word_end), so a worker that owns a part of the site axis fills the
        part of every block that lies in its words (the workers' tensors are then summed)."""
        from .sharding import shard_words
        if word_end is None:
            word_end = n_words
        if total_words is None:
            total_words = word0 + word_end
        # pf_bootstrap_counts works on groups of four elements: the per-block stride is rounded up once here,
        # so that no call ever has to pad (and copy) the whole block tensor
        elems = 3 * n_samples * n_samples
        stride = (elems + 3) // 4 * 4
        store = torch.zeros((n_blocks + n_extra_blocks, stride), dtype=torch.int32, device=self.device)
        blocks = store[:, :elems].view(n_blocks + n_extra_blocks, 3, n_samples, n_samples)
        for b in range(n_blocks):
            gb, ge = shard_words(total_words, n_blocks, b)
            wb, we = max(gb, word0) - word0, min(ge, word0 + word_end) - word0
            if we > wb:
                self.pair_counts(planes, n_samples, n_words, blocks[b], word_begin=wb, word_end=we, variant=variant)
        return blocks

    def bootstrap_weights(self, seed: int, replicate0: int, n_reps: int, n_blocks: int) -> torch.Tensor:
        w = torch.empty((n_reps, n_blocks), dtype=torch.int32, device=self.device)
        check(self.lib.pf_bootstrap_weights(seed, replicate0, n_reps, n_blocks, w.data_ptr(), _stream_ptr()))
        return w

    def bootstrap_counts(self, blocks: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
        """blocks [B, 3, n, n] int32, weights [R <= 8, B] int32 -> [R, 3, n, n] int32 weighted sums. Blocks that
        come from block_counts are used in place (their stride is a multiple of four elements); any other
        layout costs one padded copy."""
        n_blocks, shape = blocks.shape[0], tuple(blocks.shape[1:])
        elems = blocks[0].numel()
        stride = blocks.stride(0) if n_blocks > 1 else (elems + 3) // 4 * 4
        assert weights.is_contiguous() and weights.shape[1] == n_blocks
        if not (blocks[0].is_contiguous() and stride % 4 == 0 and stride >= elems and blocks.data_ptr() % 16 == 0
                and (n_blocks > 1 or elems % 4 == 0)):
            src = torch.zeros((n_blocks, (elems + 3) // 4 * 4), dtype=torch.int32, device=self.device)
            src[:, :elems] = blocks.reshape(n_blocks, elems)
            blocks, stride = src, src.stride(0)
        out = torch.empty((weights.shape[0], stride), dtype=torch.int32, device=self.device)
        check(self.lib.pf_bootstrap_counts(blocks.data_ptr(), n_blocks, stride, weights.data_ptr(), weights.shape[0],
                                           out.data_ptr(), _stream_ptr()))
        return out[:, :elems].view((weights.shape[0],) + shape)

    @staticmethod
    def bootstrap_block_plan(n_blocks: int, total_words: int, total_extra: int, total_cols: int, per_block_bytes: int,
                             max_block_bytes: int):
        """(blocks over the 2-bit words, blocks over the columns outside the 2-bit code): at most n_blocks in all and
        at most max_block_bytes of block matrices; the extra columns get their share of the blocks by column count.
        A function of the WHOLE matrix only, so every worker of a multi-GPU run - and a one-GPU run of the same file
        - cuts the same blocks."""
        cap = int(max(1, min(n_blocks, max_block_bytes // max(1, per_block_bytes))))
        bx = 0
        if total_extra > 0:
            bx = int(max(1, min(total_extra, cap - 1 if total_words > 0 and cap > 1 else cap,
                                round(cap * total_extra / max(1, total_cols + total_extra)))))
        b2 = int(max(1, min(cap - bx, total_words))) if total_words > 0 else 0
        return b2, bx

    def bootstrap_tree(self, planes: torch.Tensor, n_samples: int, n_words: int, names, replicates: int = 100,
                       seed: int = 1, n_blocks: int = 256, metric="ibs", variant: int = 0,
                       n_sites: int | None = None, max_block_bytes: int = 48 << 30,
                       extra_chars: torch.Tensor | None = None, word0: int = 0, total_cols: int | None = None,
                       extra0: int = 0, total_extra: int | None = None):
        """NJ tree of all sites with block-bootstrap support on its internal edges.
        Returns (TreeResult of the main tree with support labels in .newick, support [n-3] in percent,
        replicate join lists [replicates, n-2, 3] int32); workers other than rank 0 return the main tree without
        labels, an empty support array and None.

        Blocks: contiguous ranges of the global 32-site word axis, plus (extra_chars: the columns that do not fit
        the 2-bit code, site-major characters) contiguous ranges of those columns, in proportion to their number.
        Several workers (worker threads or a torch.distributed group; word0 / total_cols / extra0 / total_extra place this worker's
        words and extra columns in the whole matrix, phyloforge_b200/multigpu.py): every worker fills its part of
        the block matrices, one allreduce sums them, the replicates are dealt out to the workers in batches of 8
        and their join lists gathered - the draws are a function of (seed, replicate), so the support values do
        not depend on the number of workers."""
        from .bootstrap import split_support
        from .multigpu import comm_rank_world
        from .sharding import allreduce_counts, shard_words
        rank, world = comm_rank_world()
        multi = world > 1
        word_end = n_words if n_sites is None else (n_sites + 31) // 32
        local_cols = word_end * 32 if n_sites is None else n_sites
        if total_cols is None:
            total_cols = word0 * 32 + local_cols
        total_words = (total_cols + 31) // 32
        n_extra = 0 if extra_chars is None else int(extra_chars.shape[0])
        if total_extra is None:
            total_extra = extra0 + n_extra
        b2, bx = self.bootstrap_block_plan(n_blocks, total_words, total_extra, total_cols, 12 * n_samples * n_samples,
                                           max_block_bytes)
        blocks = self.block_counts(planes, n_samples, n_words, b2, word_end=word_end, variant=variant, word0=word0,
                                   total_words=total_words, n_extra_blocks=bx)
        for e in range(bx):
            gb, ge = shard_words(total_extra, bx, e)            # the same contiguous cut, on the extra-column axis
            lo, hi = max(gb, extra0) - extra0, min(ge, extra0 + n_extra) - extra0
            if hi > lo:
                self.add_extra_column_counts(extra_chars[lo:hi], n_samples, blocks[b2 + e])
        n_all = b2 + bx
        if multi:
            allreduce_counts(blocks if blocks.is_contiguous() else blocks.as_strided(
                (n_all, blocks.stride(0)), (blocks.stride(0), 1)))
        ones = torch.ones((1, n_all), dtype=torch.int32, device=self.device)
        total = self.bootstrap_counts(blocks, ones)[0].contiguous()
        main = self.tree_from_counts(total, names, metric=metric, keep=True)
        if n_samples < 4 or replicates <= 0:
            return main, np.zeros((0,)), np.zeros((0, max(0, n_samples - 2), 3), dtype=np.int32)
        rep_merges = torch.zeros((replicates, n_samples - 2, 3), dtype=torch.int32, device=self.device)
        # replicates are dealt out in eights (what one pf_bootstrap_counts call sums); several eights are joined by
        # one pf_nj_batch call, as many as fill the device: below 1,536 samples the full-scan kernel shares the SMs
        # among up to 18 matrices per launch (16 here), above it every matrix takes one 16-CTA cluster and nine
        # clusters run at a time (72 = 8 full rounds of 9; eights alone leave one slot of nine idle)
        mine = [r0 for i, r0 in enumerate(range(0, replicates, 8)) if i % world == rank]
        eights = 2 if n_samples < 1536 else (9 if 72 * 20 * n_samples * n_samples <= (16 << 30) else 1)
        for g0 in range(0, len(mine), eights):
            group = mine[g0:g0 + eights]
            dists = []
            for r0 in group:
                k = min(8, replicates - r0)
                w = self.bootstrap_weights(seed, r0, k, n_all)
                rep = self.bootstrap_counts(blocks, w)
                dists.extend(self.normalise(rep[j], metric)[0] for j in range(k))
            merge, _ = self.nj_batch(torch.stack(dists))       # the replicates of a group are joined concurrently
            at = 0
            for r0 in group:
                k = min(8, replicates - r0)
                rep_merges[r0:r0 + k] = merge[at:at + k]
                at += k
        if multi:
            allreduce_counts(rep_merges)             # every replicate was written by exactly one worker
        if rank != 0:
            # the host side (which splits of the main tree each replicate has) is rank 0's, who writes the tree: worker
            # threads of one process would only queue for the interpreter lock to compute the same numbers eight times
            return main, np.zeros((0,)), None
        rep_merges = rep_merges.cpu().numpy()
        support = split_support(main.merge, rep_merges, n_samples)
        main.newick = merges_to_newick_native(self.lib, main.merge, main.length, list(names), support=support)
        return main, support, rep_merges

    def allreduce_counts(self, counts: torch.Tensor):
        """Sum the site shards' count matrices over all ranks (NCCL over NVLink)."""
        from .sharding import allreduce_counts
        allreduce_counts(counts)

    def normalise(self, counts: torch.Tensor, metric="ibs"):
        n = counts.shape[1]
        dist_m = torch.empty((n, n), dtype=torch.float64, device=self.device)
        nz = torch.zeros(1, dtype=torch.int32, device=self.device)
        check(self.lib.pf_normalise(counts.data_ptr(), n, counts.stride(1), METRICS[metric], dist_m.data_ptr(),
                                    dist_m.stride(0), nz.data_ptr(), _stream_ptr()))
        return dist_m, nz

    def nj(self, dist_m: torch.Tensor, destroy: bool = False):
        """fp64 [n, n] -> (merge int32 [n-2, 3], length fp64 [n-2, 3]) device tensors."""
        n = dist_m.shape[0]
        if n < 3:
            raise PhyloForgeError("pf_nj needs at least 3 samples")
        work_d = dist_m if destroy else dist_m.clone()
        merge = torch.zeros((n - 2, 3), dtype=torch.int32, device=self.device)
        length = torch.zeros((n - 2, 3), dtype=torch.float64, device=self.device)
        work = torch.empty(int(self.lib.pf_nj_work_bytes(n)), dtype=torch.uint8, device=self.device)
        check(self.lib.pf_nj(work_d.data_ptr(), n, work_d.stride(0), merge.data_ptr(), length.data_ptr(),
                             work.data_ptr(), _stream_ptr()))
        return merge, length

    def nj_batch(self, dist_b: torch.Tensor):
        """fp64 [k, n, n] (overwritten) -> (merge int32 [k, n-2, 3], length fp64 [k, n-2, 3]): k trees
        joined concurrently, each by its own group of CTAs."""
        k, n = dist_b.shape[0], dist_b.shape[1]
        assert dist_b.dtype == torch.float64 and dist_b.is_contiguous() and n >= 3
        merge = torch.zeros((k, n - 2, 3), dtype=torch.int32, device=self.device)
        length = torch.zeros((k, n - 2, 3), dtype=torch.float64, device=self.device)
        work = torch.empty(k * int(self.lib.pf_nj_work_bytes(n)), dtype=torch.uint8, device=self.device)
        check(self.lib.pf_nj_batch(dist_b.data_ptr(), n, n, k, n * n, merge.data_ptr(), length.data_ptr(),
                                   work.data_ptr(), _stream_ptr()))
        return merge, length

    # ------------------------------------------------------------------ composed path
    GENOTYPE_CHARS = b"ACGTMRWSYKacgt"      # what treetool/vcftree.py:26-65 writes for a genotype; '-', 'N' (and '.') are missing

    def add_extra_column_counts(self, chars_rows: torch.Tensor, n_samples: int, counts: torch.Tensor,
                                allele_sharing: bool = True) -> int:
        """counts += the V / D1 (/ D2) of alignment columns that do not fit the 2-bit code (chars_rows: uint8
        [k, >= n_samples], one row per column, on the device): D1 = characters differ, D1 + D2 = 2 - alleles shared
        (include/phyloforge_b200.h: pf_chars_state_codes, pf_chars_allele_codes). One pack + pair-count pass per
        character state present and, with allele_sharing, per allele (A, C, G, T, *); without it D2 is left
        alone and the columns only enter the p-distance. Returns the number of states."""
        k = int(chars_rows.shape[0])
        if k == 0:
            return 0
        assert chars_rows.dtype == torch.uint8 and chars_rows.stride(1) == 1 and chars_rows.shape[1] >= n_samples
        present = [int(c) for c in torch.unique(chars_rows[:, :n_samples]).tolist()]
        states = [c for c in present if c not in (ord("-"), ord("N"), ord("."))]
        if allele_sharing and any(c not in self.GENOTYPE_CHARS for c in states):
            bad = "".join(chr(c) for c in states if c not in self.GENOTYPE_CHARS)
            raise PhyloForgeError(f"characters {bad!r} stand for no genotype: such columns can only enter the p-distance")
        if not states:
            return 0
        ldc = (n_samples + 3) // 4 * 4
        codes = torch.empty((k, ldc), dtype=torch.uint8, device=self.device)
        nw = self.n_words(k)
        planes = self.alloc_planes(n_samples, nw)
        per_state = torch.zeros_like(counts)
        for st in states:
            check(self.lib.pf_chars_state_codes(chars_rows.data_ptr(), chars_rows.stride(0), None, k, n_samples, st,
                                                codes.data_ptr(), ldc, _stream_ptr()))
            self.pack_codes(codes, n_samples, planes, nw, n_sites=k)
            self.pair_counts(planes, n_samples, nw, per_state)
        per_allele = None
        if allele_sharing:
            per_allele = torch.zeros_like(counts)
            for allele in range(5):
                check(self.lib.pf_chars_allele_codes(chars_rows.data_ptr(), chars_rows.stride(0), None, k, n_samples,
                                                     allele, codes.data_ptr(), ldc, _stream_ptr()))
                self.pack_codes(codes, n_samples, planes, nw, n_sites=k)
                self.pair_counts(planes, n_samples, nw, per_allele)
        check(self.lib.pf_counts_add_extra(per_state.data_ptr(), per_state.stride(1), len(states),
                                           per_allele.data_ptr() if per_allele is not None else None,
                                           per_allele.stride(1) if per_allele is not None else 0,
                                           counts.data_ptr(), n_samples, counts.stride(1), _stream_ptr()))
        return len(states)

    def tree_from_planes(self, planes: torch.Tensor, n_samples: int, n_words: int, names,
                         metric="ibs", variant: int = 0, keep: bool = True,
                         n_sites: int | None = None, extra_chars: torch.Tensor | None = None) -> TreeResult:
        """counts -> (allreduce) -> distances -> NJ -> Newick, on the current stream. `n_words` is
        the planes buffer's words per tile; only the words that hold the `n_sites` packed sites are
        counted (a buffer sized for an upper bound has unwritten words behind them)."""
        word_end = n_words if n_sites is None else (n_sites + 31) // 32
        counts = self.alloc_counts(n_samples)
        if word_end > 0:
            self.pair_counts(planes, n_samples, n_words, counts, word_end=word_end, variant=variant)
        if extra_chars is not None:
            self.add_extra_column_counts(extra_chars, n_samples, counts)
        self.allreduce_counts(counts)
        return self.tree_from_counts(counts, names, metric=metric, keep=keep)

    def tree_from_counts(self, counts: torch.Tensor, names, metric="ibs", keep: bool = True,
                         fetch_dist: bool = False) -> TreeResult:
        n = counts.shape[1]
        dist_m, nz = self.normalise(counts, metric)
        dist_host = None
        fetched = None
        if fetch_dist:
            # device -> pinned host buffer on the copy stream, overlapping the NJ kernel. The array
            # returned is a view of that buffer: valid until the next call that fetches distances.
            pinned = getattr(self, "_dist_pinned", None)
            if pinned is None or pinned.shape[0] != n:
                pinned = self._dist_pinned = torch.empty((n, n), dtype=torch.float64, pin_memory=True)
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream())
            copy = self._copy_stream()
            with torch.cuda.stream(copy):
                copy.wait_event(ready)
                pinned.copy_(dist_m, non_blocking=True)
                fetched = torch.cuda.Event()
                fetched.record(copy)
            dist_host = pinned.numpy()
        if n <= 2 and fetched is not None:
            fetched.synchronize()
        if n == 1:
            return TreeResult(f"{names[0]};", np.zeros((0, 3), np.int32), np.zeros((0, 3)), dist_m, counts,
                              dist_host=dist_host)
        if n == 2:
            d = float(dist_m[0, 1].item())
            return TreeResult(two_leaf_newick(names, d), np.zeros((0, 3), np.int32), np.zeros((0, 3)),
                              dist_m, counts, int(nz.item()), dist_host=dist_host)
        if fetched is not None and not keep:
            torch.cuda.current_stream().wait_event(fetched)     # NJ overwrites the matrix in place
        merge, length = self.nj(dist_m, destroy=not keep)
        merge_h = merge.cpu().numpy()
        length_h = length.cpu().numpy()
        newick = merges_to_newick_native(self.lib, merge_h, length_h, list(names))
        if fetched is not None:
            fetched.synchronize()
        return TreeResult(newick, merge_h, length_h, dist_m if keep else None, counts if keep else None,
                          int(nz.item()), dist_host=dist_host)

    def tree_from_host_codes(self, host_codes: torch.Tensor, n_samples: int, names, metric="ibs",
                             chunk_sites: int = 1 << 19, variant: int = 0, packed2: bool = False,
                             plink: bool = False) -> TreeResult:
        """End-to-end entry for genotypes that live in HOST memory: `host_codes` is a (pinned)
        site-major u8 matrix holding this rank's site shard — one byte per genotype [n_sites, ld], or,
        with packed2=True, four genotypes per byte [n_sites, ld2] (plink=True: PLINK .bed codes).
        Chunks of sites are copied host->device on a copy stream while the previous chunk is packed
        and counted; the count matrices are then allreduced over the ranks, normalised and joined,
        and the result (distance matrix, join list -> Newick) is read back to the host."""
        assert host_codes.dtype == torch.uint8 and not host_codes.is_cuda and host_codes.stride(1) == 1
        chunk_sites = max(32, chunk_sites // 32 * 32)
        m, ld = host_codes.shape[0], host_codes.stride(0)
        compute = torch.cuda.current_stream()
        copy = self._copy_stream()
        bufs = self._host_path_buffers(chunk_sites, ld, n_samples)
        counts = self.alloc_counts(n_samples)
        nw_chunk = chunk_sites // 32
        done = [None, None]
        h2d_first = h2d_last = None
        for k, c0 in enumerate(range(0, m, chunk_sites)):
            c1 = min(m, c0 + chunk_sites)
            buf = bufs["codes"][k & 1]
            with torch.cuda.stream(copy):
                if done[k & 1] is not None:
                    copy.wait_event(done[k & 1])          # the kernels that read this buffer are finished
                if h2d_first is None:
                    h2d_first = torch.cuda.Event(enable_timing=True)
                    h2d_first.record(copy)
                buf[:c1 - c0].copy_(host_codes[c0:c1], non_blocking=True)
                arrived = h2d_last = torch.cuda.Event(enable_timing=True)
                arrived.record(copy)
            compute.wait_event(arrived)
            if packed2:
                self.pack_codes_2bit(buf, n_samples, bufs["planes"], nw_chunk, n_sites=c1 - c0, plink=plink)
            else:
                self.pack_codes(buf, n_samples, bufs["planes"], nw_chunk, n_sites=c1 - c0)
            self.pair_counts(bufs["planes"], n_samples, nw_chunk, counts, word_end=(c1 - c0 + 31) // 32,
                             variant=variant)
            ev = torch.cuda.Event()
            ev.record(compute)
            done[k & 1] = ev
        self.allreduce_counts(counts)
        res = self.tree_from_counts(counts, names, metric=metric, keep=True, fetch_dist=True)
        res.d2h_bytes = int(res.dist_host.nbytes + res.merge.nbytes + res.length.nbytes)
        if h2d_first is not None:
            res.timings_ms["h2d"] = h2d_first.elapsed_time(h2d_last)     # first copy issued -> last copy landed
        return res

    # ------------------------------------------------------------------ host memory
    def numa_node(self) -> int:
        """NUMA node of this GPU's PCIe device from sysfs, or -1 when the platform does not say."""
        try:
            bus = torch.cuda.get_device_properties(self.device).pci_bus_id
            dom = torch.cuda.get_device_properties(self.device).pci_domain_id
            dev_id = torch.cuda.get_device_properties(self.device).pci_device_id
            path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev_id:02x}.0/numa_node"
            with open(path) as f:
                return int(f.read().strip())
        except (OSError, ValueError, AttributeError):
            return -1

    def pinned_host_buffer(self, shape, dtype=torch.uint8) -> torch.Tensor:
        """Page-locked host staging buffer for tree_from_host_codes, placed on the NUMA node of this GPU where the
        platform reports one (set_mempolicy(MPOL_PREFERRED) around the allocation: with one process per GPU,
        eight ranks pulling from one node's memory was what capped the 8-GPU end-to-end path in round 1)."""
        node = self.numa_node()
        libc = None
        if node >= 0:
            try:
                libc = C.CDLL(None, use_errno=True)
                mask = C.c_ulong(1 << node)
                if libc.syscall(C.c_long(238), C.c_long(1), C.byref(mask), C.c_ulong(8 * C.sizeof(C.c_ulong))) != 0:   # set_mempolicy(MPOL_PREFERRED)
                    libc = None
            except (OSError, AttributeError):
                libc = None
        try:
            buf = torch.empty(shape, dtype=dtype, pin_memory=True)       # (page-locked directly: no pageable copy first)
        finally:
            if libc is not None:
                libc.syscall(C.c_long(238), C.c_long(0), None, C.c_ulong(0))                                            # MPOL_DEFAULT
        return buf

    def release_scratch(self) -> None:
        """Drop the cached scratch buffers (tensor-core operand streams, host-path staging, pinned result)."""
        for name in ("_tc_work", "_hp", "_hp_key", "_dist_pinned", "_tc_last"):
            if hasattr(self, name):
                setattr(self, name, None)

    def _copy_stream(self):
        if getattr(self, "_copy", None) is None:
            self._copy = torch.cuda.Stream(device=self.device)
        return self._copy

    def _host_path_buffers(self, chunk_sites: int, ld: int, n_samples: int):
        key = (chunk_sites, ld, n_samples)
        if getattr(self, "_hp_key", None) != key:
            self._hp = {"codes": [torch.empty((chunk_sites, ld), dtype=torch.uint8, device=self.device)
                                  for _ in range(2)],
                        "planes": self.alloc_planes(n_samples, chunk_sites // 32)}
            self._hp_key = key
        return self._hp
