"""Host-side logic that needs no GPU: Newick assembly, config handling of the drop-in
entry points, sharding, and the world-size-2 (gloo) allreduce of site-shard counts."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from oracle import cpu as oracle, treecmp
from phyloforge_b200 import newick, phylip, run, sharding
from phyloforge_b200.script import RunCmd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_merges_to_newick_small():
    merge = np.array([[0, 1, 0], [2, 4, 3]], dtype=np.int32)
    length = np.array([[1.0, 2.0, 0.0], [3.0, 0.5, 4.0]])
    s = newick.merges_to_newick(merge, length, ["a", "b", "c", "d"])
    assert s == "(c:3,(a:1,b:2):0.5,d:4);"


def test_newick_deep_caterpillar_and_quoting():
    n = 5000
    merge = np.zeros((n - 2, 3), dtype=np.int32)
    length = np.ones((n - 2, 3))
    merge[0] = (0, 1, 0)
    for t in range(1, n - 3):
        merge[t] = (n + t - 1, t + 1, 0)
    merge[n - 3] = (n + n - 4, n - 2, n - 1)
    names = [f"s {i}" if i == 3 else f"s{i}" for i in range(n)]
    s = newick.merges_to_newick(merge, length, names)
    tree = treecmp.parse_newick(s)
    assert sorted(treecmp.leaves(tree)) == sorted(names)
    assert "'s 3'" in s
    assert newick.two_leaf_newick(["a", "b"], 0.5) == "(a:0.25,b:0.25);"


def test_phylip_distance_roundtrip(tmp_path):
    d = np.array([[0, 0.25, 1], [0.25, 0, 0.125], [1, 0.125, 0]])
    p = tmp_path / "x.dist"
    phylip.write_distance_matrix(str(p), ["a", "b", "c"], d)
    names, back = phylip.read_distance_matrix(str(p))
    assert names == ["a", "b", "c"] and np.array_equal(back, d)


@pytest.mark.parametrize("n_words,world", [(0, 1), (10, 1), (10, 3), (7, 8), (312500, 8)])
def test_shard_words_partition(n_words, world):
    spans = [sharding.shard_words(n_words, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n_words
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [e - b for b, e in spans]
    assert max(sizes) - min(sizes) <= 1


def test_shard_sites_are_word_aligned():
    spans = [sharding.shard_sites(1000, 3, r) for r in range(3)]
    assert spans == [(0, 352), (352, 704), (704, 1000)]
    with pytest.raises(ValueError):
        sharding.shard_words(10, 2, 2)


def _cfg(tmp_path, section, body):
    p = tmp_path / "run.cfg"
    p.write_text(f"[{section}]\n" + textwrap.dedent(body))
    return str(p)


def test_get_parser_matches_reference_contract(tmp_path):
    assert run.get_parser(["-s", "a.cfg"]) == ["a.cfg", "snp"]
    assert run.get_parser(["--sv", "b.cfg"]) == ["b.cfg", "sv"]
    assert run.get_parser(["-l", "c.cfg"]) == ["c.cfg", "lcn"]
    with pytest.raises(SystemExit):
        run.get_parser([])
    with pytest.raises(SystemExit):
        run.get_parser(["-v"])


def test_print_config_template(capsys):
    with pytest.raises(SystemExit):
        run.get_parser(["-c"])
    out = capsys.readouterr().out
    assert "[snp_opt]" in out and "[sv_opt]" in out and "vcf_file" in out and "tree_software" in out


def test_runcmd_reads_reference_keys(tmp_path):
    out = tmp_path / "o"
    cfg = _cfg(tmp_path, "snp_opt", f"""\
        vcf_file = /data/in.vcf
        out_path = {out}
        tree_software = TreeBest
        thread = 1
        """)
    r = RunCmd(cfg, "snp")
    assert r.vcf_file == "/data/in.vcf" and r.tree_software == "treebest"      # '*software*' values lower-cased
    assert r.thread == 2                                                       # clamped to >= 2 (script.py:69-71)
    assert r.out_path == os.path.abspath(out) and os.path.isdir(out)
    # new optional keys default to what keeps a PhyloForge config running unchanged: the treebest slot's metric,
    # multi-allelic sites kept (the reference converts them), one GPU, no bootstrap
    assert r.distance == "p" and r.non_biallelic == "include" and r.gpus == 1 and r.bootstrap == 0
    assert (r.rank, r.world) == (0, 1)
    assert r.treebest == "treebest" and r.iqtree == "iqtree"                   # software.ini


def test_runcmd_defaults_and_validation(tmp_path):
    cfg = _cfg(tmp_path, "sv_opt", f"vcf_file = x.vcf\nout_path = {tmp_path}/o\nthread = 8\n")
    assert RunCmd(cfg, "sv").tree_software == "iqtree"                         # script.py:50-52
    cfg = _cfg(tmp_path, "snp_opt", f"vcf_file = x.vcf\nout_path = {tmp_path}/o\nthread = 8\n")
    assert RunCmd(cfg, "snp").tree_software == "raxml"                         # script.py:23
    for body in ("thread = ten\n", "thread = 4\ntree_software = mrbayes\n", "thread = 4\ndistance = jc69\n",
                 "thread = 4\ngpus = 0\n", "thread = 4\ngpus = two\n", "thread = 4\nnon_biallelic = maybe\n"):
        cfg = _cfg(tmp_path, "snp_opt", f"vcf_file = x.vcf\nout_path = {tmp_path}/o\n" + body)
        with pytest.raises(SystemExit):
            RunCmd(cfg, "snp")
    with pytest.raises(SystemExit) as ei:
        RunCmd(str(tmp_path / "missing.cfg"), "snp")
    assert ei.value.code == 1


def test_treetool_namespace_is_the_drop_in():
    import treetool.run
    import treetool.script
    import treetool.svtree
    import treetool.vcftree
    assert treetool.vcftree.VcfTree.snp_tree and treetool.svtree.SvTree.sv_tree
    assert treetool.run.main is run.main and treetool.script.RunCmd is RunCmd
    for name in ("generate_ambiguous_code", "cutFile", "convert_vcf_to_seq", "merge_seq", "change_file_extension"):
        assert hasattr(treetool.vcftree.VcfTree, name)


def test_vcftree_host_helpers(tmp_path):
    from phyloforge_b200.vcftree import VcfTree
    cfg = _cfg(tmp_path, "snp_opt", f"vcf_file = x.vcf\nout_path = {tmp_path}/o\ntree_software = treebest\nthread = 3\n")
    t = VcfTree(cfg)
    assert os.path.isdir(t.working_dir)
    table = {"A": "AMRWaN", "C": "MCSYcN", "G": "RSGKgN", "T": "WYKTtN", "*": "acgtNN", ".": "NNNNN-"}
    for b1, row in table.items():
        for b2, want in zip("ACGT*.", row):
            assert t.generate_ambiguous_code(b1, b2) == want
    assert t.generate_ambiguous_code("a", "g") == "R"
    # cutFile: n // num + 1 lines per chunk, header repeated (vcftree.py:77-93)
    vcf = tmp_path / "in.vcf"
    vcf.write_text("##x\n#CHROM\tP\tI\tR\tA\tQ\tF\tI\tF\ts1\n" + "".join(f"c\t{i}\t.\tA\tG\t.\t.\t.\tGT\t0/1\n" for i in range(10)))
    parts = t.cutFile(str(vcf), 3, t.working_dir)
    assert [os.path.basename(p) for p in parts] == ["in0.vcf", "in1.vcf", "in2.vcf"]
    assert [len(open(p).read().splitlines()) - 1 for p in parts] == [4, 4, 2]
    assert t.change_file_extension(parts, ".fa")[0].endswith("in0.fa")
    for k, p in enumerate(parts):
        open(p.replace(".vcf", ".fa"), "w").write(f">s1\n{'ACG'[k] * 2}\n>s2\nNN\n")
    t.merge_seq(t.change_file_extension(parts, ".fa"), str(tmp_path / "all"), "fa")
    assert (tmp_path / "all.fa").read_text() == ">s1\nAACCGG\n>s2\nNNNNNN\n"
    t.merge_seq(t.change_file_extension(parts, ".fa"), str(tmp_path / "all"), "phylip")
    assert (tmp_path / "all.phy").read_text() == "2 6\ns1  AACCGG\ns2  NNNNNN\n"


GLOO_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from oracle import cpu as oracle
from phyloforge_b200 import sharding
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
n, m = 37, 1000
full = np.ascontiguousarray(oracle.synth_codes(n, 0, m, seed=9, miss_ppm=70000).T)
b, e = sharding.shard_sites(m, 2, rank)
part = oracle.pair_counts(np.ascontiguousarray(full[:, b:e])).astype(np.int32)
t = torch.from_numpy(part)
sharding.allreduce_counts(t)
want = oracle.pair_counts(full).astype(np.int32)
assert np.array_equal(t.numpy(), want), "allreduced shard counts differ from the unsharded counts"
assert (e - b) % 32 == 0 or e == m
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_site_shards_allreduce_world_size_2_gloo(tmp_path):
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


def test_synthetic_vcf_writers_agree_with_restated_converters():
    from oracle import ref_convert as rc
    from phyloforge_b200 import synth
    n, m = 7, 300
    names = [f"s{i}" for i in range(n)]
    codes = oracle.synth_codes(n, 0, m, seed=3, miss_ppm=100000)
    text, ref, alt = synth.snp_vcf_bytes(codes, names)
    _, seqs, kept = rc.snp_convert(text.decode())
    want = synth.expected_snp_chars(codes, ref, alt).T
    assert [s.encode() for s in seqs] == [r.tobytes() for r in want]
    assert np.array_equal(rc.snp_codes(seqs, kept)[0], codes.T)
    text, sv = synth.sv_vcf_bytes(codes, names, other_type_every=13)
    _, seqs = rc.sv_convert(text.decode())
    assert np.array_equal(rc.sv_codes(seqs), synth.expected_sv_columns(codes, sv))


def test_native_newick_serialiser_matches_the_python_one():
    """pf_newick_from_merges (host-side C++ in the library, no GPU needed) == merges_to_newick."""
    import time
    from phyloforge_b200 import _lib
    from phyloforge_b200.newick import merges_to_newick, merges_to_newick_native
    from oracle import cpu as oracle
    lib = _lib.load()
    rng = np.random.default_rng(17)
    for n in (3, 4, 5, 17, 400):
        a = rng.random((n, n))
        d = (a + a.T) / 2 - 0.05                       # some negative branch lengths
        np.fill_diagonal(d, 0.0)
        m, l = oracle.nj(d)
        names = [f"s{i}" for i in range(n)]
        names[0] = "odd name(1)"
        names[-1] = "it's:quoted"
        if n > 4:
            names[2] = "ünï"
        assert merges_to_newick_native(lib, m, l, names) == merges_to_newick(m, l, names)
        if n > 3:
            sup = rng.integers(0, 101, size=n - 3).astype(np.float64)
            sup[0] = 12.5
            assert merges_to_newick_native(lib, m, l, names, support=sup) == merges_to_newick(m, l, names, support=sup)
    # caterpillar: depth ~ n, must stay linear
    n = 3000
    pos = np.cumsum(np.arange(1, n + 1, dtype=np.float64) * 0.01)
    d = np.abs(pos[:, None] - pos[None, :]) + 1.0
    np.fill_diagonal(d, 0.0)
    m, l = oracle.nj(d)
    names = [f"S{i:05d}" for i in range(n)]
    t0 = time.perf_counter()
    text = merges_to_newick_native(lib, m, l, names)
    t1 = time.perf_counter()
    assert text == merges_to_newick(m, l, names) and t1 - t0 < 0.5
    bad = m.copy()
    bad[0, 0] = n + 5                                   # refers to a node that does not exist yet
    with pytest.raises(ValueError):
        merges_to_newick_native(lib, bad, l, names)


def test_tensor_path_workspace_size_of_degenerate_ranges():
    """pf_pair_counts_tc_work_bytes is pure host arithmetic: empty and tiny word ranges must not divide by zero
    (found by tools/fuzz_gpu.py: an empty range raised SIGFPE)."""
    from phyloforge_b200 import _lib
    lib = _lib.load()
    assert lib.pf_pair_counts_tc_work_bytes(300, 0) > 0
    assert lib.pf_pair_counts_tc_work_bytes(0, 10) > 0
    sizes = [lib.pf_pair_counts_tc_work_bytes(300, w) for w in (1, 2, 15, 16, 17, 63, 64, 65, 1000)]
    assert all(b > 0 for b in sizes) and sizes == sorted(sizes)
    assert lib.pf_pair_counts_tc_work_bytes(10000, 156250) > 2 * 25_000_000_000 // 2      # two operand streams at C5


def test_bgzf_reader_inflates_blocks_in_order(tmp_path):
    """phyloforge_b200.vcf._BgzfReader (parallel inflate of bgzip blocks) returns the original bytes for any read
    size, recognises BGZF by its 'BC' subfield (a plain gzip file is left to the gzip module), and reports damage."""
    import gzip
    import numpy as np
    import vcfgen
    from phyloforge_b200 import vcf
    rng = np.random.default_rng(3)
    data = bytes(rng.integers(32, 127, size=700_001, dtype=np.uint8)) + b"\n"
    path = tmp_path / "x.vcf.gz"
    vcfgen.write_bgzf(path, data, block=40_000, seed=1)
    assert gzip.open(path, "rb").read() == data                      # a valid multi-member gzip file
    assert isinstance(vcf._open_stream(str(path)), vcf._BgzfReader)
    for size in (1, 4096, 65536 * 3 + 17, len(data) + 5):
        out = bytearray()
        with vcf._BgzfReader(str(path), threads=3) as r:
            buf = bytearray(size)
            while True:
                k = r.readinto(buf)
                if not k:
                    break
                out += buf[:k]
                if len(out) > len(data):
                    break
        assert bytes(out) == data, size
    plain = tmp_path / "p.vcf.gz"
    with gzip.open(plain, "wb") as f:
        f.write(data)
    assert not isinstance(vcf._open_stream(str(plain)), vcf._BgzfReader)
    raw = bytearray(path.read_bytes())
    broken = tmp_path / "b.vcf.gz"
    broken.write_bytes(bytes(raw[:len(raw) // 2]))                  # cut in the middle of a block
    with pytest.raises(vcf.VcfFormatError):
        with vcf._BgzfReader(str(broken)) as r:
            buf = bytearray(1 << 20)
            while r.readinto(buf):
                pass
    raw[300] ^= 0x55                                                 # flip bits inside the first block's deflate stream
    broken.write_bytes(bytes(raw))
    with pytest.raises(vcf.VcfFormatError):
        with vcf._BgzfReader(str(broken)) as r:
            buf = bytearray(1 << 20)
            while r.readinto(buf):
                pass


@pytest.mark.parametrize("decimals", [10, 6, 0])
def test_native_distance_writer_matches_python_formatting(tmp_path, decimals):
    """pf_write_distance_matrix (host threads, exact 128-bit fixed-point formatting) against Python's "%.<d>f": ties at
    the last digit, rationals k / V as the path produces them, zeros, tiny and large values, NaN, padded rows."""
    rng = np.random.default_rng(decimals)
    n = 211
    vals = [0.0, 1.0, 0.5, 0.25, 0.125, 2.0 ** -33, 2.0 ** -34, 2.0 ** -35, 3 * 2.0 ** -35, 5e-11, 1.5e-10, 2.5e-10,
            0.99999999995, 0.999999999949999, 1023.99999999995, 1e-300, 5e-324, 1024.0, 123456.789, 1e22, -0.25,
            float("nan"), 0.30000000000000004, 1 / 3, 2 / 3, 0.1, 0.7]
    vals += [k / (1 << 11) + 2.0 ** -12 for k in range(40)]                       # exact ties for few decimals
    vals += [(2 * k + 1) / 2 ** (decimals + 1) / 5.0 ** decimals for k in range(40)]   # close to ...5 at the last digit
    v = rng.integers(1, 10_000_000, size=n * n)
    k = (rng.random(n * n) * v).astype(np.int64)
    d = (k / v).reshape(n, n)
    flat = d.reshape(-1)
    flat[:len(vals)] = vals
    ld = n + 5
    padded = np.full((n, ld), 7.0)
    padded[:, :n] = d
    names = [f"s{i}" for i in range(n)]
    names[3] = "with blank"
    want = tmp_path / "want.dist"
    phylip.write_distance_matrix_py(str(want), names, d, fmt=f"%.{decimals}f")
    for threads in (1, 3, 0):
        got = tmp_path / f"got{threads}.dist"
        phylip.write_distance_matrix(str(got), names, padded[:, :n] if threads else d, decimals=decimals, threads=threads)
        assert got.read_bytes() == want.read_bytes(), threads
    # empty and one-sample matrices; an unwritable path fails loudly
    phylip.write_distance_matrix(str(tmp_path / "e.dist"), [], np.zeros((0, 0)))
    assert (tmp_path / "e.dist").read_bytes() == b"0\n"
    phylip.write_distance_matrix(str(tmp_path / "o.dist"), ["x"], np.zeros((1, 1)), decimals=2)
    assert (tmp_path / "o.dist").read_bytes() == b"1\nx  0.00\n"
    from phyloforge_b200._lib import PhyloForgeError
    with pytest.raises(PhyloForgeError):
        phylip.write_distance_matrix(str(tmp_path / "no" / "such" / "dir.dist"), ["x"], np.zeros((1, 1)))
