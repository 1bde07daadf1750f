"""Host logic of `gpus = N` (phyloforge_b200/multigpu.py, vcf.byte_range_of_rank) that needs no GPU: the byte
ranges the workers convert tile the VCF body at line starts (the reference's axis: contiguous line ranges in file
order, treetool/vcftree.py:77-93), the ranged chunk reader hands out exactly those lines, chunk FASTAs are merged in
worker order like merge_seq does (vcftree.py:145-171), the bootstrap block plan does not depend on the number of
workers, the workers' process management, and the world-size-2 (gloo) exchange helpers."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from oracle import ref_convert as rc
from phyloforge_b200 import multigpu, sharding, vcf
from phyloforge_b200.engine import Engine

import vcfgen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _body_lines(path):
    with open(path, "rb") as f:
        data = f.read()
    names, size, body = vcf.read_header(path, with_body_offset=True)
    return data, body


def _read_range(path, rng, buf_bytes):
    reader = vcf._ChunkReader(path, rng[1] - rng[0], buf_bytes, byte_range=rng)
    out = []
    try:
        for bi, nb in reader.chunks():
            out.append(bytes(reader.bufs[bi].numpy()[:nb]))
            reader.release(bi)
    finally:
        reader.close()
    return b"".join(out)


@pytest.mark.parametrize("case", ["snp_rand_b.vcf", "snp_rand_crlf.vcf", "sv_rand_a.vcf", "snp_edge.vcf"])
@pytest.mark.parametrize("world", [1, 2, 3, 8, 64])
def test_byte_ranges_tile_the_body_at_line_starts(golden_dir, case, world):
    path = os.path.join(golden_dir, case)
    data, body = _body_lines(path)
    ranges = [vcf.byte_range_of_rank(path, body, r, world) for r in range(world)]
    assert ranges[0][0] == body and ranges[-1][1] == len(data)
    for (b0, e0), (b1, e1) in zip(ranges, ranges[1:]):
        assert e0 == b1 and b0 <= e0                       # contiguous, in file order, possibly empty
    for b, e in ranges:
        assert b == body or b == len(data) or data[b - 1:b] == b"\n"     # every cut is a line start
    # the ranged reader hands out exactly the worker's lines, whatever the chunk size
    for buf in (4096, 1 << 20):
        longest = max(len(ln) for ln in data.split(b"\n")) + 2
        got = b"".join(_read_range(path, rng, max(buf, 2 * longest)) for rng in ranges)
        assert got == data[body:]


def test_byte_ranges_without_trailing_newline(tmp_path):
    text, _ = vcfgen.random_snp_vcf(3, 5, 40)
    p = tmp_path / "x.vcf"
    p.write_text(text.rstrip("\n"))
    data, body = _body_lines(str(p))
    for world in (2, 5):
        ranges = [vcf.byte_range_of_rank(str(p), body, r, world) for r in range(world)]
        assert b"".join(data[b:e] for b, e in ranges) == data[body:]


def test_workers_lines_convert_to_the_reference_columns_in_order(golden_dir):
    """Converting every worker's lines separately (with the restated reference converter) and concatenating per sample
    gives the whole-file result: cutting at line starts commutes with the conversion."""
    path = os.path.join(golden_dir, "snp_rand_b.vcf")
    data, body = _body_lines(path)
    header = data[:body].decode()
    names, seqs, _ = rc.snp_convert(data.decode())
    for world in (2, 5):
        parts = []
        for r in range(world):
            b, e = vcf.byte_range_of_rank(path, body, r, world)
            parts.append(rc.snp_convert(header + data[b:e].decode())[1])
        assert ["".join(p[i] for p in parts) for i in range(len(names))] == list(seqs)


def test_merge_chunk_fastas_is_merge_seq(tmp_path):
    rng = np.random.default_rng(0)
    names = [f"S{i}" for i in range(7)]
    blocks = [rng.choice(np.frombuffer(b"ACGT-N", dtype=np.uint8), size=(7, w)) for w in (5, 0, 13, 1)]
    paths = []
    for k, blk in enumerate(blocks):
        p = tmp_path / f"in{k}.fa"
        vcf.write_fasta(str(p), names, blk)
        paths.append(str(p))
    whole = np.concatenate(blocks, axis=1)
    out = multigpu.merge_chunk_fastas(paths, str(tmp_path / "all_snp"), "fa")
    vcf.write_fasta(str(tmp_path / "want.fa"), names, whole)
    assert open(out, "rb").read() == open(tmp_path / "want.fa", "rb").read()
    out = multigpu.merge_chunk_fastas(paths, str(tmp_path / "all_snp"), "phylip")
    vcf.write_phylip(str(tmp_path / "want.phy"), names, whole)
    assert open(out, "rb").read() == open(tmp_path / "want.phy", "rb").read()
    # worker threads hand the shards themselves to the writers (no chunk file is read back): the same bytes
    vcf.write_fasta(str(tmp_path / "parts.fa"), names, blocks)
    vcf.write_phylip(str(tmp_path / "parts.phy"), names, [b[:, ::1] for b in blocks])
    assert open(tmp_path / "parts.fa", "rb").read() == open(tmp_path / "want.fa", "rb").read()
    assert open(tmp_path / "parts.phy", "rb").read() == open(tmp_path / "want.phy", "rb").read()


def test_bootstrap_block_plan_depends_on_the_whole_matrix_only():
    plan = Engine.bootstrap_block_plan
    assert plan(256, 10_000, 0, 320_000, 12 * 50 * 50, 48 << 30) == (256, 0)
    assert plan(256, 100, 0, 3200, 12 * 50 * 50, 48 << 30) == (100, 0)            # never more blocks than words
    b2, bx = plan(256, 10_000, 16_000, 320_000, 12 * 50 * 50, 48 << 30)
    assert b2 + bx == 256 and bx == round(256 * 16_000 / 336_000)
    assert plan(256, 10_000, 3, 320_000, 12 * 50 * 50, 48 << 30) == (255, 1)      # a few extra columns: one block
    assert plan(256, 0, 40, 0, 12 * 50 * 50, 48 << 30) == (0, 40)                 # only extra columns
    assert plan(256, 10_000, 0, 320_000, 12 * 3000 * 3000, 2 << 30) == (19, 0)    # memory cap


def test_worker_env_and_spawn(tmp_path, monkeypatch):
    assert multigpu.worker_env() == (0, 1)
    monkeypatch.setenv(multigpu.ENV_RANK, "3")
    monkeypatch.setenv(multigpu.ENV_WORLD, "8")
    assert multigpu.worker_env() == (3, 8)
    monkeypatch.setenv(multigpu.ENV_RANK, "9")
    assert multigpu.worker_env() == (0, 1)
    monkeypatch.delenv(multigpu.ENV_RANK)
    monkeypatch.delenv(multigpu.ENV_WORLD)
    assert multigpu.spawn_workers(["-v"], 3) == 0                               # every worker prints the version
    assert multigpu.spawn_workers(["-s", str(tmp_path / "missing.cfg")], 2) == 1  # a failing worker fails the run


WORKER = """
import sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from phyloforge_b200 import multigpu
rank = int(sys.argv[1])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=2)
assert multigpu.gather_ints([10 + rank, 7]) == [[10, 7], [11, 7]]
assert multigpu.agree_ok(True) is True
assert multigpu.agree_ok(rank == 0) is False
multigpu.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_exchange_helpers_world_size_2_gloo(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


def test_planes_shift_restatement():
    """numpy restatement of pf_planes_shift on the tiled layout (what the GPU test compares the kernel with)."""
    from oracle import cpu as oracle
    rng = np.random.default_rng(1)
    n, m, shift = 70, 77, 13
    codes = rng.integers(0, 4, size=(n, m), dtype=np.uint8)
    padded = np.concatenate([np.full((n, shift), 3, np.uint8), codes], axis=1)
    want = oracle.pack_planes(padded)
    got = oracle.shift_planes(oracle.pack_planes(codes), n, m, shift)
    assert np.array_equal(got, want)


# ---------------------------------------------------------------- worker threads (gpu_workers = threads)
def _threads(world, body):
    """Run body(rank) on `world` worker threads of a ThreadGroup; returns (status, per-rank results)."""
    results = [None] * world

    def work():
        rank, w = multigpu.worker_env()
        assert w == world
        results[rank] = body(rank)
        return 0

    return multigpu.run_threads(work, world), results


@pytest.mark.parametrize("world", [2, 3, 8])
def test_worker_threads_exchange_helpers(world):
    def body(rank):
        assert multigpu.comm_rank_world() == (rank, world)
        table = multigpu.gather_ints([10 + rank, 7])
        assert table == [[10 + r, 7] for r in range(world)]
        assert multigpu.agree_ok(True) is True
        assert multigpu.agree_ok(rank != world - 1) is False
        multigpu.barrier()
        return "ok"

    status, results = _threads(world, body)
    assert status == 0 and results == ["ok"] * world
    assert multigpu.worker_env() == (0, 1) and multigpu.comm_rank_world() == (0, 1)     # the caller is no worker


@pytest.mark.parametrize("world,shape", [(2, (3, 5, 5)), (3, (3, 7, 7)), (8, (3, 2, 2)), (4, (5,)), (5, (2, 3))])
def test_worker_threads_sum_counts_through_slices(world, shape):
    """thread_allreduce_sum without NCCL (host tensors here; device tensors of workers that share a GPU in the -m gpu
    tests): every worker ends with the integer sum, also when there are more workers than elements per slice."""
    import torch
    rng = np.random.default_rng(world)
    parts = [rng.integers(-1000, 1000, size=shape).astype(np.int32) for _ in range(world)]
    want = sum(p.astype(np.int64) for p in parts).astype(np.int32)

    def body(rank):
        t = torch.from_numpy(parts[rank].copy())
        out = sharding.allreduce_counts(t)
        assert out is t
        again = torch.from_numpy(parts[rank].copy())           # a second call in the same group
        sharding.allreduce_counts(again)
        return t.numpy(), again.numpy()

    status, results = _threads(world, body)
    assert status == 0
    for got, again in results:
        assert np.array_equal(got, want) and np.array_equal(again, want)


def test_worker_threads_sum_rows_with_padding():
    """Block matrices with a padded row stride (engine.bootstrap_tree) are cut along the first axis."""
    import torch
    world = 3
    base = [torch.arange(4 * 10, dtype=torch.int32).reshape(4, 10) * (r + 1) for r in range(world)]

    def body(rank):
        view = base[rank].as_strided((4, 7), (10, 1))
        assert not view.is_contiguous()
        sharding.allreduce_counts(view)
        return base[rank].clone()

    status, results = _threads(world, body)
    assert status == 0
    want = torch.arange(40, dtype=torch.int32).reshape(4, 10) * 6
    for r, got in enumerate(results):
        assert torch.equal(got[:, :7], want[:, :7])
        assert torch.equal(got[:, 7:], (torch.arange(40, dtype=torch.int32).reshape(4, 10) * (r + 1))[:, 7:])   # padding untouched


def test_a_failing_worker_thread_fails_the_run_and_releases_the_others():
    def work():
        rank, _ = multigpu.worker_env()
        if rank == 1:
            sys.exit(3)                       # the workflow's own exit (a record the reference cannot convert, ...)
        multigpu.barrier()                    # the others would wait here forever
        return 0

    assert multigpu.run_threads(work, 3) == 3

    def work2():
        if multigpu.worker_env()[0] == 0:
            raise RuntimeError("boom")
        multigpu.gather_ints([1])
        return 0

    assert multigpu.run_threads(work2, 2) == 1
    assert multigpu.run_threads(lambda: None, 2) == 0      # a workflow that returns nothing succeeded


def test_gpu_workers_config_key(tmp_path, capsys):
    from phyloforge_b200.script import RunCmd
    src = tmp_path / "x.vcf"
    src.write_text("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tA\tB\n")
    cfg = tmp_path / "run.cfg"
    base = f"[snp_opt]\nvcf_file = {src}\nout_path = {tmp_path / 'out'}\ntree_software = treebest\nthread = 4\n"
    cfg.write_text(base)
    assert RunCmd(str(cfg), "snp").gpu_workers == "threads"            # default: worker threads of one process
    cfg.write_text(base + "gpus = 2\ngpu_workers = Processes\n")
    rc = RunCmd(str(cfg), "snp")
    assert (rc.gpus, rc.gpu_workers) == (2, "processes")
    cfg.write_text(base + "gpu_workers = fibers\n")
    with pytest.raises(SystemExit):
        RunCmd(str(cfg), "snp")
    assert "gpu_workers" in capsys.readouterr().out


@pytest.mark.parametrize("threads", [1, 3, 8])
def test_chunk_reader_with_several_positional_reads_per_buffer(golden_dir, monkeypatch, threads):
    """A buffer of a plain-text file is filled by several reads in flight at once (pieces of READ_PIECE bytes): the
    chunks handed out are the same whole lines, for the whole file and for a worker's byte range."""
    monkeypatch.setattr(vcf._ChunkReader, "READ_PIECE", 1000)
    path = os.path.join(golden_dir, "snp_rand_b.vcf")
    data, body = _body_lines(path)
    longest = max(len(ln) for ln in data.split(b"\n")) + 2

    def read(rng, buf):
        reader = vcf._ChunkReader(path, len(data) if rng is None else rng[1] - rng[0], buf, byte_range=rng,
                                  read_threads=threads)
        out = []
        try:
            for bi, nb in reader.chunks():
                chunk = bytes(reader.bufs[bi].numpy()[:nb])
                assert chunk.endswith(b"\n") or len(b"".join(out)) + len(chunk) == len(data) - (0 if rng is None else rng[0])
                out.append(chunk)
                reader.release(bi)
        finally:
            reader.close()
        return b"".join(out)

    for buf in (2 * longest, 7919, 1 << 20):
        buf = max(buf, 2 * longest)
        assert read(None, buf) == data
        for world in (2, 3):
            got = b"".join(read(vcf.byte_range_of_rank(path, body, r, world), buf) for r in range(world))
            assert got == data[body:]


def test_a_meeting_everybody_reached_stays_reached():
    """The worker that holds the failure message learns from agree_ok that the run failed, prints, and exits; a
    faster peer that exits first (and aborts the group) must not take that answer away from it."""
    for _ in range(200):
        said = []

        def work():
            rank, _ = multigpu.worker_env()
            ok = multigpu.agree_ok(rank != 1)
            if rank == 1:
                said.append(ok)               # (the failure text of the real workflow)
            sys.exit(1)

        assert multigpu.run_threads(work, 4) == 1
        assert said == [False]


@pytest.mark.parametrize("threads", [1, 2, 8])
def test_alignment_files_written_by_positional_writes(tmp_path, monkeypatch, threads):
    """Large alignment files are written by several threads at the rows' offsets (vcf._write_rows): the same bytes as
    the sequential writer, for one matrix, for column blocks (the workers' shards), for rows that are views of a wider
    array (the loader's result), for empty blocks and for a single sample."""
    monkeypatch.setattr(vcf, "PARALLEL_WRITE_MIN_BYTES", 1)
    rng = np.random.default_rng(threads)
    n = 37
    names = [f"sample_{i}" + "x" * (i % 5) for i in range(n)]           # headers of different lengths
    wide = rng.choice(np.frombuffer(b"ACGTMRWSYK-N", dtype=np.uint8), size=(n, 5000))
    blocks = [wide[:, :1234], wide[:, 1234:1234], np.ascontiguousarray(wide[:, 1234:4000]), wide[:, 4000:]]
    whole = np.ascontiguousarray(wide)

    def sequential(path, first, heads):
        with open(path, "wb") as f:
            f.write(first)
            for h, row in zip(heads, whole):
                f.write(h + row.tobytes() + b"\n")

    sequential(tmp_path / "want.fa", b"", [b">" + x.encode() + b"\n" for x in names])
    sequential(tmp_path / "want.phy", f"{n} {whole.shape[1]}\n".encode(), [x.encode() + b"  " for x in names])
    for tag, chars in (("one", whole), ("blocks", blocks), ("view", wide[:, :5000])):
        vcf.write_fasta(str(tmp_path / f"{tag}.fa"), names, chars, threads=threads)
        vcf.write_phylip(str(tmp_path / f"{tag}.phy"), names, chars, threads=threads)
        assert (tmp_path / f"{tag}.fa").read_bytes() == (tmp_path / "want.fa").read_bytes(), tag
        assert (tmp_path / f"{tag}.phy").read_bytes() == (tmp_path / "want.phy").read_bytes(), tag
    vcf.write_fasta(str(tmp_path / "single.fa"), names[:1], whole[:1], threads=threads)
    assert (tmp_path / "single.fa").read_bytes() == b">" + names[0].encode() + b"\n" + whole[0].tobytes() + b"\n"
    vcf.write_fasta(str(tmp_path / "none.fa"), names, np.zeros((n, 0), dtype=np.uint8), threads=threads)
    assert (tmp_path / "none.fa").read_bytes() == b"".join(b">" + x.encode() + b"\n\n" for x in names)
    # rewriting a longer file in place leaves no tail of the old one
    vcf.write_fasta(str(tmp_path / "one.fa"), names[:3], whole[:3, :10], threads=threads)
    assert (tmp_path / "one.fa").read_bytes() == b"".join(b">" + x.encode() + b"\n" + r.tobytes() + b"\n" for x, r in zip(names[:3], whole[:3, :10]))


def _bgzf_lines(path, rank, world, threads=2):
    with vcf._BgzfLines(str(path), rank, world, threads=threads) as r:
        out = bytearray()
        buf = bytearray(1237)                          # an awkward read size
        while True:
            k = r.readinto(buf)
            if k == 0:
                return bytes(out)
            out += buf[:k]


@pytest.mark.parametrize("block", [40, 700, 5000, 0xff00])
@pytest.mark.parametrize("world", [2, 3, 8, 33])
def test_bgzf_file_cut_over_workers_hands_out_every_line_once(golden_dir, tmp_path, block, world):
    """A bgzip file is cut at block starts; every worker but the first drops the text through its first newline, every
    worker but the last reads on through the next one: the workers' texts, in rank order, are the file's text, each
    made of whole lines - for blocks much shorter and much longer than a line, and for more workers than blocks."""
    text = open(os.path.join(golden_dir, "snp_rand_b.vcf"), "rb").read()
    path = tmp_path / "x.vcf.gz"
    vcfgen.write_bgzf(path, text, block=block, seed=block + world)
    size = os.path.getsize(path)
    cuts = [vcf.bgzf_range_of_rank(str(path), r, world) for r in range(world)]
    assert cuts[0][0] == 0 and cuts[-1][1] == size
    for (b0, e0), (b1, e1) in zip(cuts, cuts[1:]):
        assert e0 == b1 and b0 <= e0
    raw = open(path, "rb").read()
    for b, _ in cuts:
        assert b == size or vcf._bgzf_block_size(raw[b:b + 18 + 64]) >= 26          # every cut is a block start
    parts = [_bgzf_lines(path, r, world) for r in range(world)]
    assert b"".join(parts) == text
    for part in parts[:-1]:
        assert part == b"" or part.endswith(b"\n")
    # and through the chunk reader, as the loader reads it
    got = b""
    for r in range(world):
        reader = vcf._ChunkReader(str(path), None, 1 << 16, bgzf_share=(r, world))
        try:
            for bi, nb in reader.chunks():
                got += bytes(reader.bufs[bi].numpy()[:nb])
                reader.release(bi)
        finally:
            reader.close()
    assert got == text


def test_bgzf_cut_corner_cases(tmp_path):
    # no newline at the end; one very long line over many blocks; deflate payloads that contain the gzip magic
    rng = np.random.default_rng(5)
    magic = b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00"
    long_line = bytes(rng.integers(65, 91, size=30_000, dtype=np.uint8))
    noise = [bytes(rng.integers(0, 256, size=int(k), dtype=np.uint8)).replace(b"\n", b"x") + magic for k in rng.integers(1, 300, size=400)]
    text = b"#h1\n#h2\n" + long_line + b"\n" + b"\n".join(noise) + b"\nlast line without newline"
    for block, world in ((64, 5), (3000, 4), (50_000, 3), (50_000, 16)):
        path = tmp_path / f"c{block}_{world}.gz"
        vcfgen.write_bgzf(path, text, block=block, seed=block)
        parts = [_bgzf_lines(path, r, world) for r in range(world)]
        assert b"".join(parts) == text, (block, world)
        for r, p_r in enumerate(parts[:-1]):          # whole lines; the unterminated last line has one owner
            assert p_r == b"" or p_r.endswith(b"\n") or all(q == b"" for q in parts[r + 1:])
    # a gzip file that bgzip did not write cannot be entered in the middle
    import gzip
    plain = tmp_path / "plain.vcf.gz"
    with gzip.open(plain, "wb") as f:
        f.write(text)
    with pytest.raises(vcf.VcfFormatError, match="bgzip"):
        vcf.bgzf_range_of_rank(str(plain), 1, 2)
