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
import io
import os
import queue
import threading
from concurrent.futures import ThreadPoolExecutor
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
# first guess of the text size of a gzip / bgzip file (ratio x compressed size + slack): its planes start at that
# size (at most 256 MB) and double whenever the packed columns outgrow them
GZ_TEXT_RATIO_GUESS = 16
GZ_TEXT_SLACK = 1 << 20


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
    extra_chars: torch.Tensor | None = None   # uint8 [n_not_2bit, ld] on the device (non_biallelic = include): the
                                              # characters of those sites, site-major, for Engine.add_extra_column_counts


def _is_gzip(path: str) -> bool:
    with open(path, "rb") as f:
        return f.read(2) == b"\x1f\x8b"


def _open_bytes(path: str):
    """Binary reader of the VCF text. gzip / bgzip files (``.vcf.gz``; not accepted by the reference,
    SURVEY.md §8(f).1) are inflated on the host while the previous chunk is parsed on the GPU."""
    if _is_gzip(path):
        import gzip
        return gzip.open(path, "rb")
    return open(path, "rb", buffering=0)


def _bgzf_block_size(header: bytes) -> int:
    """Total size of the BGZF block that starts with `header` (>= 18 bytes), or 0 if this gzip member is not a
    BGZF block (SAM specification section 4.1: a gzip member with FEXTRA and the 'BC' subfield = BSIZE - 1)."""
    if len(header) < 18 or header[:4] != b"\x1f\x8b\x08\x04":
        return 0
    xlen = int.from_bytes(header[10:12], "little")
    extra = header[12:12 + xlen]
    pos = 0
    while pos + 4 <= len(extra):
        slen = int.from_bytes(extra[pos + 2:pos + 4], "little")
        if extra[pos:pos + 2] == b"BC" and slen == 2:
            return int.from_bytes(extra[pos + 4:pos + 6], "little") + 1
        pos += 4 + slen
    return 0


class _BgzfReader:
    """`readinto` over a bgzip file (what `bgzip` / htslib write: independent gzip members of <= 64 KB). The file is
    read in large pieces, cut at block boundaries, and groups of blocks are inflated (and CRC-checked) by a pool of
    host threads a few groups ahead - zlib releases the GIL - and handed out in file order; one Python gzip stream
    inflates ~0.2 GB/s, which is what `.vcf.gz` input was bound by."""

    GROUP_BYTES = 2 << 20        # compressed bytes per task
    READ_BYTES = 8 << 20

    def __init__(self, path: str, threads: int | None = None, c_begin: int = 0, c_end: int | None = None):
        """c_begin / c_end: offsets of block starts in the compressed file; the blocks in [c_begin, c_end) are handed
        out by readinto, `read_through_newline` then goes on behind c_end (a worker's share of a file cut over
        several workers, _BgzfLines)."""
        from collections import deque
        self.path = path
        self.f = open(path, "rb", buffering=0)
        self.f.seek(c_begin)
        self.c_pos = c_begin         # offset in the file of self.raw[0]
        self.c_end = c_end
        self.threads = threads or max(2, min(16, (os.cpu_count() or 4)))
        self.pool = ThreadPoolExecutor(self.threads)
        self.ahead = 2 * self.threads
        self.pending = deque()
        self.raw = b""               # compressed bytes read but not yet cut into tasks
        self.cur = memoryview(b"")
        self.raw_eof = False

    @staticmethod
    def _inflate_group(path: str, raw: bytes, blocks):
        import zlib
        parts = []
        for off, size, xlen in blocks:
            data = zlib.decompress(raw[off + 12 + xlen:off + size - 8], -15)
            crc = int.from_bytes(raw[off + size - 8:off + size - 4], "little")
            isize = int.from_bytes(raw[off + size - 4:off + size], "little")
            if len(data) != isize or zlib.crc32(data) != crc:
                raise VcfFormatError(f"{path}: corrupt BGZF block (length or CRC32 does not match its trailer)")
            parts.append(data)
        return b"".join(parts)

    def _submit_next(self) -> bool:
        """Cut one group of whole blocks off the front of the compressed bytes and hand it to the pool."""
        while True:
            blocks, off = [], 0
            ended = False                # the next block starts at or behind c_end: not this reader's
            while off + 18 <= len(self.raw) and off < self.GROUP_BYTES:
                if self.c_end is not None and self.c_pos + off >= self.c_end:
                    ended = True
                    break
                size = _bgzf_block_size(self.raw[off:off + 18 + 64])
                if size < 26:
                    raise VcfFormatError(f"{self.path}: not a BGZF block where one is expected (mixed or damaged gzip file)")
                if off + size > len(self.raw):
                    break
                blocks.append((off, size, int.from_bytes(self.raw[off + 10:off + 12], "little")))
                off += size
            if blocks and (off >= self.GROUP_BYTES or self.raw_eof or ended):
                chunk, self.raw = self.raw[:off], self.raw[off:]
                self.c_pos += off
                self.pending.append(self.pool.submit(self._inflate_group, self.path, chunk, blocks))
                return True
            if ended:
                return False
            if self.raw_eof:
                if self.raw:
                    raise VcfFormatError(f"{self.path}: truncated BGZF block")
                return False
            more = self.f.read(self.READ_BYTES)
            if not more:
                self.raw_eof = True
            else:
                self.raw += more

    def readinto(self, out) -> int:
        mv = memoryview(out).cast("B")
        filled = 0
        while filled < len(mv):
            if not len(self.cur):
                while len(self.pending) < self.ahead and self._submit_next():
                    pass
                if not self.pending:
                    break
                try:
                    self.cur = memoryview(self.pending.popleft().result())
                except VcfFormatError:
                    raise
                except Exception as e:  # zlib.error
                    raise VcfFormatError(f"{self.path}: corrupt BGZF block ({e})") from None
                continue
            k = min(len(self.cur), len(mv) - filled)
            mv[filled:filled + k] = self.cur[:k]
            self.cur = self.cur[k:]
            filled += k
        return filled

    def read_through_newline(self) -> bytes:
        """After readinto has handed out everything up to c_end: the text from there through the first newline (or
        to the end of the file), block by block."""
        out = []
        while True:
            while len(self.raw) < 18 and not self.raw_eof:
                more = self.f.read(1 << 16)
                if not more:
                    self.raw_eof = True
                else:
                    self.raw += more
            if len(self.raw) == 0:
                break
            size = _bgzf_block_size(self.raw[:18 + 64])
            if size < 26:
                raise VcfFormatError(f"{self.path}: not a BGZF block where one is expected (mixed or damaged gzip file)")
            while len(self.raw) < size and not self.raw_eof:
                more = self.f.read(1 << 16)
                if not more:
                    self.raw_eof = True
                else:
                    self.raw += more
            if len(self.raw) < size:
                raise VcfFormatError(f"{self.path}: truncated BGZF block")
            try:
                text = self._inflate_group(self.path, self.raw[:size], [(0, size, int.from_bytes(self.raw[10:12], "little"))])
            except VcfFormatError:
                raise
            except Exception as e:  # zlib.error
                raise VcfFormatError(f"{self.path}: corrupt BGZF block ({e})") from None
            self.raw = self.raw[size:]
            self.c_pos += size
            k = text.find(b"\n")
            if k >= 0:
                out.append(text[:k + 1])
                break
            out.append(text)
        return b"".join(out)

    def close(self):
        self.pool.shutdown(wait=False, cancel_futures=True)
        self.f.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def _bgzf_block_start_at_or_after(f, pos: int, size: int) -> int:
    """Offset of the first BGZF block that starts at `pos` or later (`size` when there is none). A candidate - the
    gzip magic with FEXTRA and a 'BC' subfield - counts when the two blocks behind it chain up as well (or the file
    ends there): deflate data can hold the magic bytes, three consistent headers in a row it does not."""
    if pos <= 0:
        return 0
    at = pos
    while at < size:
        f.seek(at)
        window = f.read((1 << 18) + 128)              # more than three blocks of at most 64 KB
        if len(window) < 18:
            return size
        k = window.find(b"\x1f\x8b\x08\x04")
        while k >= 0:
            ok, q = True, k
            for _ in range(3):
                if at + q == size:
                    break
                bs = _bgzf_block_size(window[q:q + 18 + 64])
                if bs < 26 or at + q + bs > size:
                    ok = False
                    break
                q += bs
                if q + 18 > len(window):
                    break                              # (a chain that leaves the window was long enough)
            if ok:
                return at + k
            k = window.find(b"\x1f\x8b\x08\x04", k + 1)
        at += max(1, len(window) - 17)
    return size


def bgzf_range_of_rank(path: str, rank: int, world: int):
    """(begin, end) offsets in a bgzip file of the blocks worker `rank` of `world` starts from: the compressed bytes are
    cut evenly and every cut is moved forward to the next block start. The lines of the text are dealt out along these
    cuts by _BgzfLines. Raises for gzip files that are not BGZF (one deflate stream cannot be entered in the middle)."""
    size = os.path.getsize(path)
    with open(path, "rb", buffering=0) as f:
        if not _bgzf_block_size(f.read(18 + 256)):
            raise VcfFormatError(f"{path}: a gzip file that was not written by bgzip cannot be cut over several workers; "
                                 "use gpus = 1 (or bgzip / unpack it)")
        cut = [_bgzf_block_start_at_or_after(f, size * r // world, size) if r else 0 for r in (rank, rank + 1)]
    if rank + 1 == world:
        cut[1] = size
    return cut[0], max(cut[1], cut[0])


class _BgzfLines:
    """`readinto` over worker `rank`'s lines of a bgzip file cut at block starts (bgzf_range_of_rank). The blocks'
    text is one stream of lines; a line belongs to the worker in whose blocks its first newline-terminated
    predecessor ends - in practice: every worker but the first drops the text through the first newline of its
    blocks, every worker but the last goes on behind its blocks through the next newline. Every line of the file is
    handed out by exactly one worker (tests/test_multigpu_host.py), whatever the block and line lengths."""

    def __init__(self, path: str, rank: int, world: int, threads: int | None = None):
        c_begin, c_end = bgzf_range_of_rank(path, rank, world)
        self.inner = _BgzfReader(path, threads=threads, c_begin=c_begin, c_end=c_end)
        self.skipping = rank > 0          # still looking for the newline behind which this worker's lines start
        self.owns_tail = rank + 1 < world
        self.started = rank == 0          # a line of this worker is open or may follow
        self.tail = None                  # the text behind the blocks, once it has been read
        self.tail_at = 0

    def readinto(self, out) -> int:
        mv = memoryview(out).cast("B")
        filled = 0
        while filled < len(mv):
            if self.tail is not None:
                k = min(len(self.tail) - self.tail_at, len(mv) - filled)
                if k <= 0:
                    break
                mv[filled:filled + k] = self.tail[self.tail_at:self.tail_at + k]
                self.tail_at += k
                filled += k
                continue
            got = self.inner.readinto(mv[filled:])
            if got == 0:
                # the blocks are exhausted: a worker that never found its first newline owns nothing (the line
                # it sat in belongs to a worker before it, who reads through that line's end)
                self.tail = self.inner.read_through_newline() if (self.owns_tail and self.started) else b""
                continue
            if self.skipping:
                piece = bytes(mv[filled:filled + got])
                k = piece.find(b"\n")
                if k < 0:
                    continue                       # all of it belongs to the worker before
                self.skipping = False
                self.started = True
                rest = piece[k + 1:]
                mv[filled:filled + len(rest)] = rest
                filled += len(rest)
                continue
            filled += got
        return filled

    def close(self):
        self.inner.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def _open_stream(path: str):
    """Reader with `readinto` for the chunk reader: bgzip files through the parallel block reader, other gzip
    files through the gzip module, plain text unbuffered."""
    if _is_gzip(path):
        with open(path, "rb") as f:
            if _bgzf_block_size(f.read(18 + 256)):
                return _BgzfReader(path)
    return _open_bytes(path)


_HEADER_LINES: dict = {}     # path -> lines up to and including #CHROM (set by read_header)


def read_header(path: str, with_body_offset: bool = False):
    """Sample names from the first line that starts with ``#CHROM`` (vcftree.py:69-73,
    99-100). Returns (names, size in bytes of the text; None for gzip input, whose text size is not known
    up front) and, with_body_offset, the byte offset just behind the ``#CHROM`` line."""
    size = os.path.getsize(path)
    if _is_gzip(path):
        size = None
    names = None
    body = n_header = 0
    with _open_bytes(path) as f:
        for raw in f:
            body += len(raw)
            n_header += 1
            _HEADER_LINES[path] = n_header
            if raw.startswith(b"#CHROM"):
                names = raw.decode("utf-8", "replace").strip().split()[9:]
                break
            if not raw.startswith(b"#") and raw.strip():
                break
    if names is None:
        raise VcfFormatError(f"{path}: no #CHROM header line before the first record")
    if not names:
        raise VcfFormatError(f"{path}: the #CHROM header names no samples")
    if with_body_offset:
        return names, size, body
    return names, size


def line_start_at_or_after(f, pos: int, size: int) -> int:
    """Offset of the first line that starts at byte `pos` or later (a line starts at 0 or just behind a newline)."""
    if pos <= 0:
        return 0
    if pos >= size:
        return size
    f.seek(pos - 1)
    at = pos - 1
    while at < size:
        block = f.read(1 << 16)
        if not block:
            break
        k = block.find(b"\n")
        if k >= 0:
            return at + k + 1
        at += len(block)
    return size


def byte_range_of_rank(path: str, body_offset: int, rank: int, world: int):
    """The contiguous run of whole lines of the VCF body that worker `rank` of `world` converts: the body bytes are
    cut evenly and every cut is moved forward to the next line start - the axis the reference cuts over its worker
    processes (contiguous line ranges in file order, treetool/vcftree.py:77-93). Returns (begin, end) byte offsets;
    the ranges of ranks 0..world-1 tile [body_offset, file size) exactly."""
    size = os.path.getsize(path)
    span = max(0, size - body_offset)
    with open(path, "rb", buffering=0) as f:
        cut = [line_start_at_or_after(f, body_offset + span * r // world, size) if r else body_offset
               for r in (rank, rank + 1)]
    if rank + 1 == world:
        cut[1] = size
    return max(cut[0], body_offset), max(cut[1], cut[0], body_offset)


def _pread_full(fd: int, out, offset: int) -> int:
    """Positional read of len(out) bytes at `offset` into `out` (fewer only at the end of the file)."""
    total = 0
    n = len(out)
    while total < n:
        got = os.preadv(fd, [out[total:]], offset + total)
        if got <= 0:
            break
        total += got
    return total


class _ChunkReader:
    """Background reader: fills pinned host buffers with whole lines of the file. A chunk ends just
    after a newline; the partial line behind it is carried to the front of the next buffer. The
    read (page cache -> pinned memory) of chunk k+1 overlaps the GPU work on chunk k."""

    # plain-text files: a buffer is filled by several positional reads at once (os.preadv releases the interpreter
    # lock; one thread copies page cache -> pinned memory at ~5 GB/s, which was what the whole conversion waited for)
    READ_PIECE = 8 << 20
    READ_THREADS = 8

    def __init__(self, path: str, size: int, buf_bytes: int, n_bufs: int = 2, byte_range=None, read_threads=None,
                 bgzf_share=None):
        self.path, self.size, self.buf_bytes = path, size, buf_bytes
        self.byte_range = byte_range          # (begin, end) of a plain-text file, cut at line starts
        self.bgzf_share = bgzf_share          # (rank, world): this worker's lines of a bgzip file (_BgzfLines)
        if read_threads is None:
            from . import multigpu
            world = max(1, multigpu.worker_env()[1])      # worker threads / processes of one box share its cores
            read_threads = max(1, min(self.READ_THREADS, (os.cpu_count() or 1) // world))
        self.read_threads = read_threads
        # page-locked staging buffers (plain ones only where no CUDA driver exists: the reader's line cutting is
        # exercised by the CPU tests; the loader itself needs the GPU)
        pin = torch.cuda.is_available()
        self.bufs = [torch.empty(buf_bytes, dtype=torch.uint8, pin_memory=True) if pin else
                     torch.empty(buf_bytes, dtype=torch.uint8) for _ in range(n_bufs)]
        self.free = queue.Queue()
        self.full = queue.Queue()
        for i in range(n_bufs):
            self.free.put(i)
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _read_pieces(self, pool, fd: int, out, pos: int):
        """Fill `out` from file offset `pos`: cut into one piece per reader thread (>= READ_PIECE bytes each), one
        positional read per piece, all in flight together. Returns (bytes read in file order, whether the file
        ended inside `out` - it shrank under us)."""
        want = len(out)
        piece = max(self.READ_PIECE, -(-want // self.read_threads))
        cuts = list(range(0, want, piece)) + [want]
        spans = list(zip(cuts, cuts[1:]))
        jobs = [pool.submit(_pread_full, fd, out[a:b], pos + a) for a, b in spans]
        got, short = 0, False
        for (a, b), job in zip(spans, jobs):
            k_got = job.result()
            if not short:
                got += k_got
            short = short or k_got < b - a
        return got, short

    def _run(self):
        pool = None
        try:
            carry = b""
            done = 0
            eof = False
            with (_BgzfLines(self.path, *self.bgzf_share) if self.bgzf_share else _open_stream(self.path)) as f:
                left = None
                if self.byte_range is not None:
                    f.seek(self.byte_range[0])
                    left = self.byte_range[1] - self.byte_range[0]
                    eof = left <= 0
                if self.read_threads > 1 and isinstance(f, io.FileIO) and hasattr(os, "preadv"):
                    pool = ThreadPoolExecutor(self.read_threads, thread_name_prefix="phyloforge-read")
                    pos = f.tell()
                    if left is None:
                        left = max(0, os.fstat(f.fileno()).st_size - pos)
                        eof = left <= 0
                while not eof or carry:
                    i = self.free.get()
                    if i is None:
                        return
                    arr = self.bufs[i].numpy()
                    k = len(carry)
                    if k >= self.buf_bytes:
                        raise VcfFormatError(f"{self.path}: a line is longer than the {self.buf_bytes}-byte chunk buffer")
                    if k:
                        arr[:k] = np.frombuffer(carry, dtype=np.uint8)
                    filled = k
                    mv = memoryview(arr)
                    if pool is not None and not eof:
                        got, short = self._read_pieces(pool, f.fileno(), mv[filled:filled + min(left, self.buf_bytes - filled)], pos)
                        filled += got
                        done += got
                        pos += got
                        left -= got
                        eof = left <= 0 or short
                    while pool is None and filled < self.buf_bytes and not eof:
                        want = mv[filled:] if left is None else mv[filled:filled + min(left, self.buf_bytes - filled)]
                        got = f.readinto(want)
                        if not got:
                            eof = True
                            break
                        filled += got
                        done += got
                        if left is not None:
                            left -= got
                            eof = left <= 0
                    if eof:
                        cut = filled                       # last chunk: take everything
                    else:
                        tail = arr[max(0, filled - (1 << 20)):filled]
                        nl = np.flatnonzero(tail == 10)
                        if nl.size == 0:
                            nl = np.flatnonzero(arr[:filled] == 10)
                            if nl.size == 0:
                                raise VcfFormatError(f"{self.path}: a line is longer than the chunk buffer")
                            cut = int(nl[-1]) + 1
                        else:
                            cut = max(0, filled - (1 << 20)) + int(nl[-1]) + 1
                    carry = arr[cut:filled].tobytes()
                    self.full.put((i, cut))
            self.full.put(None)
        except BaseException as e:  # noqa: BLE001 - hand the failure to the consumer
            self.full.put(e)
        finally:
            if pool is not None:
                pool.shutdown(wait=False, cancel_futures=True)

    def chunks(self):
        while True:
            item = self.full.get()
            if item is None:
                return
            if isinstance(item, BaseException):
                raise item
            yield item

    def release(self, i: int):
        self.free.put(i)

    def close(self):
        self.free.put(None)


class _CharSink:
    """Where the character matrix of the alignment file (all_snp.fa / SV.matrix) goes as the chunks are converted:
    every chunk's [n, columns] block is copied device -> one of two page-locked staging buffers (asynchronously,
    behind the kernels that made it) and from there, by a host thread, into its columns of the result - one pageable
    [n, bound] array when the text size bounds the columns (its untouched tail costs nothing), else a list of
    blocks joined at the end. Page-locking a fresh buffer per chunk cost more than the copy it served."""

    COPY_THREADS = 4

    def __init__(self, n: int, cols_bound):
        self.n, self.off, self.error = n, 0, None
        self.final = None
        if cols_bound is not None:
            try:
                self.final = np.empty((n, max(1, int(cols_bound))), dtype=np.uint8)
            except MemoryError:
                self.final = None
        self.blocks = []
        self.staging = [None, None]
        self.free = [threading.Event(), threading.Event()]
        for e in self.free:
            e.set()
        self.k = 0
        self.q = queue.Queue()
        self.pool = ThreadPoolExecutor(self.COPY_THREADS, thread_name_prefix="phyloforge-chars") if n >= 64 else None
        self.thread = threading.Thread(target=self._run, name="phyloforge-chars", daemon=True)
        self.thread.start()

    def push(self, block: torch.Tensor):
        """block: device uint8 [n, c], complete on the current stream once the work queued so far has run."""
        c = int(block.shape[1])
        k = self.k
        self.k ^= 1
        self.free[k].wait()
        if self.error is not None:
            raise self.error
        self.free[k].clear()
        need = self.n * c
        if self.staging[k] is None or self.staging[k].numel() < need:
            self.staging[k] = None
            self.staging[k] = torch.empty(need + need // 8, dtype=torch.uint8, pin_memory=True)
        stage = self.staging[k][:need].view(self.n, c)
        stage.copy_(block, non_blocking=True)
        done = torch.cuda.Event()
        done.record()
        self.q.put((k, done, stage, block))        # (the device block stays alive until its copy has run)

    def _run(self):
        while True:
            item = self.q.get()
            if item is None:
                if self.pool is not None:
                    self.pool.shutdown(wait=False)
                return
            k, done, stage, _block = item
            try:
                if self.error is None:
                    done.synchronize()
                    c = stage.shape[1]
                    if self.final is not None and self.off + c <= self.final.shape[1]:
                        src, dst = stage.numpy(), self.final[:, self.off:self.off + c]
                        if self.pool is None or self.n * c < (8 << 20):
                            dst[:] = src
                        else:                            # row ranges on a few threads (numpy copies outside the lock)
                            cuts = [self.n * j // self.COPY_THREADS for j in range(self.COPY_THREADS + 1)]
                            jobs = [self.pool.submit(np.copyto, dst[a:b], src[a:b]) for a, b in zip(cuts, cuts[1:]) if b > a]
                            for job in jobs:
                                job.result()
                    else:
                        if self.final is not None:       # more columns than the bound (cannot happen for text that
                            self.blocks.append(self.final[:, :self.off].copy())   # parses): keep going in blocks
                            self.final = None
                        self.blocks.append(stage.numpy().copy())
                    self.off += c
            except BaseException as e:  # noqa: BLE001 - handed to the loader
                self.error = e
            finally:
                self.free[k].set()

    def finish(self) -> np.ndarray:
        self.q.put(None)
        self.thread.join()
        if self.error is not None:
            raise self.error
        if self.final is not None:
            return self.final[:, :self.off]
        if not self.blocks:
            return np.zeros((self.n, 0), dtype=np.uint8)
        return self.blocks[0] if len(self.blocks) == 1 else np.concatenate(self.blocks, axis=1)

    def abandon(self):
        if self.thread.is_alive():
            self.q.put(None)


def _load(engine, path: str, kind: str, keep_chars: bool, non_biallelic: str, chunk_bytes: int,
          rank: int = 0, world: int = 1):
    lib, dev = engine.lib, engine.device
    names, size, body = read_header(path, with_body_offset=True)
    n = len(names)
    rpl = 1 if kind == "snp" else 2          # matrix rows a line can yield
    ld = (n + 3) // 4 * 4
    stream = torch.cuda.current_stream().cuda_stream
    if size == 0:
        raise VcfFormatError(f"{path}: empty file")
    byte_range = bgzf_share = None
    if world > 1:
        if size is None:
            bgzf_range_of_rank(path, rank, world)     # (raises for gzip files that bgzip did not write)
            bgzf_share = (rank, world)                # cut at block starts; the header lines are some worker's lines too
        else:
            byte_range = byte_range_of_rank(path, body, rank, world)
            size = byte_range[1] - byte_range[0]

    # A record that yields a matrix column has 9 fixed fields and n sample fields of >= 1 byte, each
    # followed by a separator: at least 2n + 18 bytes. That bounds the columns of a plain-text file and
    # (for well-formed input) the lines of a chunk; blank-line-heavy chunks fall back to counting. The text
    # size of a gzip file is not known up front: its planes start small and grow geometrically.
    min_line = 2 * n + 18
    known = size is not None
    buf_bytes = int(min(max(chunk_bytes, 4096), size + 1)) if known else int(max(chunk_bytes, 4096))
    text_bound = size if known else GZ_TEXT_RATIO_GUESS * os.path.getsize(path) + GZ_TEXT_SLACK
    n_words_cap = max(1, ((text_bound // min_line + 2) * rpl + 31) // 32)
    if not known:
        n_words_cap = max(1, min(n_words_cap, (1 << 28) // max(1, 512 * ((n + 63) // 64))))   # <= 256 MB to begin with
    planes = engine.alloc_planes(n, n_words_cap)
    max_lines = buf_bytes // min_line + 2

    def grow_planes(need_words):
        """Planes of a larger capacity holding the words packed so far (tiles keep their [word][plane][lane] order)."""
        nonlocal planes, n_words_cap
        new_cap = max(need_words, 2 * n_words_cap)
        bigger = engine.alloc_planes(n, new_cap)
        n_tiles = (n + 63) // 64
        bigger.view(n_tiles, new_cap, 128)[:, :words_done] = planes.view(n_tiles, n_words_cap, 128)[:, :words_done]
        planes, n_words_cap = bigger, new_cap

    if known and size <= 0:
        # a worker whose byte range holds no line: an empty shard (all of its one word is code 3)
        planes.fill_(-1)
        return PackedGenotypes(names=names, n_samples=n, n_sites=0, n_words=n_words_cap, planes=planes,
                               chars=np.zeros((n, 0), dtype=np.uint8) if keep_chars else None, n_columns=0, n_lines=0)

    reader = _ChunkReader(path, size, buf_bytes, byte_range=byte_range, bgzf_share=bgzf_share)
    d_text = torch.empty(buf_bytes, dtype=torch.uint8, device=dev)
    d_work = torch.empty(int(lib.pf_line_index_work_bytes(buf_bytes)), dtype=torch.uint8, device=dev)
    carry_rows = 32     # rows [0, 32) of the code buffer carry the sites left over from the previous chunk

    def alloc_line_buffers(cap):
        return (torch.empty(cap + 1, dtype=torch.int64, device=dev),
                torch.empty(cap, dtype=torch.int32, device=dev),
                torch.empty((carry_rows + cap * rpl, ld), dtype=torch.uint8, device=dev),
                torch.empty((cap * rpl, ld), dtype=torch.uint8, device=dev),
                torch.empty(carry_rows + cap * rpl, dtype=torch.int64, device=dev),
                torch.empty(cap * rpl, dtype=torch.int64, device=dev) if (keep_chars or non_biallelic == "include") else None)

    d_lines, d_info, d_codes, d_chars, d_rows, d_rows_all = alloc_line_buffers(max_lines)
    n_carry = 0            # leftover sites (< 32) waiting in d_codes[0:n_carry]
    words_done = 0
    # lines before this worker's first line: the header for the first of several workers (so that its line
    # numbers are the file's and the workers' line counts add up to the file's), nothing otherwise
    n_lines_total = _HEADER_LINES.get(path, 0) if (world > 1 and rank == 0 and bgzf_share is None) else 0
    n_unfit_total = 0
    h2d = 0
    sink = _CharSink(n, (size // min_line + 2) * rpl if known else None) if keep_chars else None
    extra_blocks = []
    summary = (C.c_int64 * 5)()
    parse = lib.pf_vcf_snp_parse if kind == "snp" else lib.pf_vcf_sv_parse
    try:
        from .multigpu import trace
        for (bi, nb) in reader.chunks():
            trace(f"chunk of {nb} bytes read")
            h_buf = reader.bufs[bi]
            d_text[:nb].copy_(h_buf[:nb], non_blocking=True)
            h2d += nb
            n_lines = C.c_int64()
            rc = lib.pf_line_index(d_text.data_ptr(), nb, d_lines.data_ptr(), max_lines + 1, C.byref(n_lines),
                                   d_work.data_ptr(), stream)
            if rc == 3:   # PF_ERR_INPUT: more lines than the bound (blank / short lines): count, grow, retry
                carried = d_codes[:n_carry].clone()
                max_lines = int(np.count_nonzero(h_buf[:nb].numpy() == 10)) + 2
                d_lines, d_info, d_codes, d_chars, d_rows, d_rows_all = alloc_line_buffers(max_lines)
                d_codes[:n_carry] = carried
                rc = lib.pf_line_index(d_text.data_ptr(), nb, d_lines.data_ptr(), max_lines + 1, C.byref(n_lines),
                                       d_work.data_ptr(), stream)
            check(rc)
            reader.release(bi)          # pf_line_index synchronised the stream: the H2D copy is done
            trace("chunk on the device, lines indexed")
            nl = n_lines.value
            codes_body = d_codes[carry_rows:]
            check(parse(d_text.data_ptr(), nb, d_lines.data_ptr(), nl, n, d_chars.data_ptr(),
                        codes_body.data_ptr(), ld, d_info.data_ptr(), stream))
            require_fit = 1 if kind == "snp" else 0
            # gather list for packing: leftover rows first, then this chunk's surviving rows
            check(lib.pf_vcf_select_rows(d_info.data_ptr(), nl, rpl, require_fit,
                                         d_rows[carry_rows:].data_ptr(), max_lines * rpl, summary, stream))
            n_cols, n_err, n_unfit, first_line, first_cls = (int(x) for x in summary)
            if n_err:
                raise VcfFormatError(
                    f"{path}: {n_err} record(s) the reference cannot convert; first at line "
                    f"{n_lines_total + first_line + 1}: {ERR_TEXT.get(first_cls, first_cls)}")
            if n_unfit and non_biallelic not in ("skip", "include"):
                raise VcfFormatError(
                    f"{path}: {n_unfit} kept site(s) are not biallelic A/C/G/T and do not fit the 2-bit "
                    "genotype code (set `non_biallelic = skip` in the config to leave them out of the "
                    "distance matrix, or `non_biallelic = include` to count them from their characters; "
                    "all_snp.fa contains them either way)")
            n_unfit_total += n_unfit
            if keep_chars or (n_unfit and non_biallelic == "include"):
                if n_unfit:
                    check(lib.pf_vcf_select_rows(d_info.data_ptr(), nl, rpl, 0, d_rows_all.data_ptr(),
                                                 max_lines * rpl, summary, stream))
                    n_all = int(summary[0])
                    rows_all = d_rows_all[:n_all]
                else:
                    n_all = n_cols
                    rows_all = d_rows[carry_rows:carry_rows + n_cols]
            if n_unfit and non_biallelic == "include":
                # the kept rows that are not in the fit list: their characters stay on the device
                is_unfit = torch.zeros(max_lines * rpl, dtype=torch.bool, device=dev)
                is_unfit[rows_all] = True
                is_unfit[d_rows[carry_rows:carry_rows + n_cols]] = False
                extra_blocks.append(d_chars.index_select(0, is_unfit.nonzero().flatten()))
            if keep_chars:
                if n_all:
                    sink.push(engine.transpose_u8(d_chars, n, rows=rows_all, n_rows=n_all))
            # pack: full words now, remainder carried to the next chunk
            d_rows[carry_rows:carry_rows + n_cols] += carry_rows        # rows index into d_codes
            first = carry_rows - n_carry
            if n_carry:
                d_rows[first:carry_rows] = torch.arange(0, n_carry, dtype=torch.int64, device=dev)
            avail = n_carry + n_cols
            full = (avail // 32) * 32
            if words_done + (avail + 31) // 32 > n_words_cap:
                if known:
                    raise VcfFormatError(f"{path}: more matrix columns than the file size allows")
                grow_planes(words_done + (avail + 31) // 32)
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
            trace("chunk parsed, packed (queued)")
        if n_carry:
            rows = torch.arange(0, n_carry, dtype=torch.int64, device=dev)
            engine.pack_codes(d_codes, n, planes, n_words_cap, word0=words_done, rows=rows, n_sites=n_carry)
        n_sites = words_done * 32 + n_carry
        if n_sites == 0:
            planes.fill_(-1)           # a worker whose share holds no site: an empty shard (every word is code 3)
        torch.cuda.current_stream().synchronize()
        trace("device work of the conversion done")
    except BaseException:
        if sink is not None:
            sink.abandon()             # the copier thread ends behind what was queued
        raise
    finally:
        reader.close()

    chars = None
    n_columns = n_sites
    if keep_chars:
        chars = sink.finish()
        n_columns = chars.shape[1]
        trace("character matrix on the host")
    return PackedGenotypes(names=names, n_samples=n, n_sites=n_sites, n_words=n_words_cap, planes=planes,
                           chars=chars, n_columns=n_columns, n_lines=n_lines_total,
                           n_not_2bit=n_unfit_total, h2d_bytes=h2d,
                           extra_chars=torch.cat(extra_blocks, dim=0) if extra_blocks else None)


def load_snp_vcf(engine, path: str, keep_chars: bool = True, non_biallelic: str = "include",
                 chunk_bytes: int = DEFAULT_CHUNK_BYTES, rank: int = 0, world: int = 1) -> PackedGenotypes:
    """SNP VCF -> planes of the biallelic sites + the IUPAC character matrix of every kept site.
    world > 1: only the lines of worker `rank`'s byte range (byte_range_of_rank) are converted."""
    return _load(engine, path, "snp", keep_chars, non_biallelic, chunk_bytes, rank, world)


def load_sv_vcf(engine, path: str, keep_chars: bool = True,
                chunk_bytes: int = DEFAULT_CHUNK_BYTES, rank: int = 0, world: int = 1) -> PackedGenotypes:
    """SV VCF -> planes and the 0/1/. matrix (one or two columns per DEL/INS record)."""
    return _load(engine, path, "sv", keep_chars, "error", chunk_bytes, rank, world)


PARALLEL_WRITE_MIN_BYTES = 32 << 20
WRITE_THREADS = 8


def _pwritev_all(fd: int, bufs, offset: int):
    """All of `bufs` at `offset` (a positional write may stop short: very long rows, signals)."""
    bufs = [memoryview(b).cast("B") for b in bufs if len(b)]
    while bufs:
        done = os.pwritev(fd, bufs[:512], offset)
        offset += done
        while bufs and done >= len(bufs[0]):
            done -= len(bufs[0])
            bufs.pop(0)
        if bufs and done:
            bufs[0] = bufs[0][done:]


def _write_rows(path: str, first_line: bytes, heads, chars, threads=None):
    """first_line, then per sample i: heads[i], row i of every block of `chars`, a newline. chars: one [n, m]
    character matrix or a list of them (column blocks of the same samples in order: the workers' shards, concatenated
    per sample as merge_seq does with the chunk files, vcftree.py:145-157); rows must be contiguous. Large files are
    written by several threads with positional writes at the rows' offsets (one thread copies into the page cache
    at ~3 GB/s, which made the alignment file the longest step of a converted VCF)."""
    parts = [p for p in (list(chars) if isinstance(chars, (list, tuple)) else [chars])]
    parts = [p if p.ndim == 2 and (p.shape[1] == 0 or p.strides[1] == 1) else np.ascontiguousarray(p) for p in parts]
    parts = [p for p in parts if p.shape[1]]
    total = sum(int(p.shape[1]) for p in parts)
    n = len(heads)
    if threads is None:
        try:
            threads = int(os.environ.get("PHYLOFORGE_B200_WRITE_THREADS", WRITE_THREADS))
        except ValueError:
            threads = WRITE_THREADS
        threads = max(1, min(threads, os.cpu_count() or 1))
    offsets = np.zeros(n + 1, dtype=np.int64)
    offsets[0] = len(first_line)
    if n:
        offsets[1:] = len(first_line) + np.cumsum([len(h) + total + 1 for h in heads])
    size = int(offsets[n])
    if threads <= 1 or size < PARALLEL_WRITE_MIN_BYTES or n < 2 or not hasattr(os, "pwritev"):
        with open(path, "wb", buffering=1 << 20) as f:
            f.write(first_line)
            for i in range(n):
                f.write(heads[i])
                for p in parts:
                    f.write(memoryview(p[i]))
                f.write(b"\n")
        return total
    fd = os.open(path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o666)
    try:
        os.ftruncate(fd, size)
        if first_line:
            _pwritev_all(fd, [first_line], 0)

        def job(lo, hi):
            for i in range(lo, hi):
                _pwritev_all(fd, [heads[i], *[p[i] for p in parts], b"\n"], int(offsets[i]))

        n_jobs = min(n, 4 * threads)
        cuts = [n * j // n_jobs for j in range(n_jobs + 1)]
        with ThreadPoolExecutor(threads, thread_name_prefix="phyloforge-write") as pool:
            for fut in [pool.submit(job, a, b) for a, b in zip(cuts, cuts[1:]) if b > a]:
                fut.result()
    finally:
        os.close(fd)
    return total


def write_fasta(path: str, names, chars, threads=None):
    """``>id\\nseq\\n`` per sample (vcftree.py:168-171, svtree.py:122-125). chars: a matrix or a list of column
    blocks (see _write_rows)."""
    _write_rows(path, b"", [b">" + name.encode() + b"\n" for name in names], chars, threads)


def write_phylip(path: str, names, chars, threads=None):
    """``N L`` then ``id  seq`` per sample (vcftree.py:159-166)."""
    parts = list(chars) if isinstance(chars, (list, tuple)) else [chars]
    total = sum(int(p.shape[1]) for p in parts)
    _write_rows(path, f"{len(names)} {total}\n".encode(), [name.encode() + b"  " for name in names], chars, threads)
