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
