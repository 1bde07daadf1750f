"""`gpus = N` through the entry point: N workers (threads of this process by default, processes with
`gpu_workers = processes`), each converting its byte range of the VCF on its GPU, count matrices summed over the
workers, rank 0 writing the artefacts - the outputs must equal the one-GPU run byte for byte (character matrix,
distance matrix, tree with bootstrap support). On a box with fewer GPUs than workers the workers share a device
(threads sum through device-to-device copies, processes through gloo), which exercises everything but NCCL itself;
with `gpurun --gpus 2` the same tests run over NCCL (in-process for threads, a process group for processes)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import cpu as oracle
from phyloforge_b200 import _lib, run

import vcfgen

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,m,shift", [(70, 77, 13), (64, 32, 31), (130, 1000, 1), (5, 64, 0), (200, 33, 17)])
def test_planes_shift_matches_restatement(engine, n, m, shift):
    rng = np.random.default_rng(n + m + shift)
    codes = rng.integers(0, 4, size=(n, m), dtype=np.uint8)
    src = torch.from_numpy(oracle.pack_planes(codes).view(np.int32)).to(engine.device)
    nw_src, nw_dst = (m + 31) // 32, (m + shift + 31) // 32
    dst = engine.alloc_planes(n, nw_dst + 2)          # a roomier destination: words behind the last one stay untouched
    dst.fill_(0x5a5a5a5a)
    _lib.check(engine.lib.pf_planes_shift(src.data_ptr(), nw_src, dst.data_ptr(), nw_dst + 2, n, m, shift,
                                          torch.cuda.current_stream().cuda_stream))
    got = dst.cpu().numpy().view(np.uint32).reshape(-1, nw_dst + 2, 128)
    want = oracle.shift_planes(oracle.pack_planes(codes), n, m, shift).reshape(-1, nw_dst, 128)
    assert np.array_equal(got[:, :nw_dst], want)
    assert (got[:, nw_dst:] == 0x5a5a5a5a).all()
    # and the shifted matrix unpacks to the padded codes
    back = engine.unpack_codes(dst, n, nw_dst + 2, m + shift).cpu().numpy()
    assert (back[:, :shift] == 3).all() and np.array_equal(back[:, shift:], codes)


def _cfg(path, section, vcf_file, out, tool, extra=""):
    path.write_text(f"[{section}]\nvcf_file = {vcf_file}\nout_path = {out}\ntree_software = {tool}\nthread = 4\n{extra}")


def _run(tmp_path, tag, section, flag, src, tool, extra):
    out = tmp_path / f"out_{tag}"
    cfg = tmp_path / f"run_{tag}.cfg"
    _cfg(cfg, section, src, out, tool, extra)
    run.main([flag, str(cfg)])
    return out


@pytest.mark.parametrize("gpus,workers", [(2, "threads"), (3, "threads"), (8, "threads"), (2, "processes")])
def test_snp_entry_point_with_several_workers_equals_one_gpu(engine, tmp_path, gpus, workers):
    """multi-allelic / '*' sites included, both metrics' ingredients, seeded bootstrap: every artefact identical."""
    text, _ = vcfgen.random_snp_vcf(17, 41, 2600, miss=0.04, multi_p=0.15, star_p=0.04)
    src = tmp_path / "in.vcf"
    src.write_text(text)
    extra = "distance = ibs\nbootstrap = 24\nbootstrap_seed = 5\n"
    one = _run(tmp_path, "1", "snp_opt", "-s", src, "treebest", extra)
    many = _run(tmp_path, str(gpus), "snp_opt", "-s", src, "treebest",
                extra + f"gpus = {gpus}\ngpu_workers = {workers}\n")
    for name in ("all_snp.fa", "SNP.dist", "SNP_treebest.NHX"):
        assert (one / name).read_bytes() == (many / name).read_bytes(), name
    rep1 = json.loads((one / "SNP.phyloforge_b200.json").read_text())
    repn = json.loads((many / "SNP.phyloforge_b200.json").read_text())
    assert repn["gpus"] == gpus and rep1["gpus"] == 1
    for key in ("samples", "sites", "lines", "alignment_columns", "sites_not_2bit", "bootstrap_mean_support"):
        assert rep1[key] == repn[key], key
    # worker processes exchange through the reference's layout, one chunk FASTA per worker in working_dir; worker
    # threads hand their shards over in memory and leave working_dir empty, as one GPU does
    assert sorted(os.listdir(many / "working_dir")) == ([f"in{r}.fa" for r in range(gpus)] if workers == "processes" else [])


def test_snp_phylip_for_external_tools_with_two_workers(engine, golden_dir, tmp_path):
    src = os.path.join(golden_dir, "snp_rand_a.vcf")
    out = tmp_path / "out"
    cfg = tmp_path / "run.cfg"
    _cfg(cfg, "snp_opt", src, out, "iqtree", "gpus = 2\n")
    with pytest.raises(SystemExit) as ei:            # iqtree is absent: the child's exit status, like the reference
        run.main(["-s", str(cfg)])
    assert ei.value.code not in (0, None)
    assert (out / "all_snp.phy").read_bytes() == open(os.path.join(golden_dir, "snp_rand_a.all_snp.phy"), "rb").read()


@pytest.mark.parametrize("workers", ["threads", "processes"])
def test_sv_entry_point_with_two_workers_equals_one_gpu(engine, golden_dir, tmp_path, workers):
    src = os.path.join(golden_dir, "sv_rand_b.vcf")
    one = _run(tmp_path, "1", "sv_opt", "-S", src, "nj", "bootstrap = 8\n")
    two = _run(tmp_path, "2", "sv_opt", "-S", src, "nj", f"bootstrap = 8\ngpus = 2\ngpu_workers = {workers}\n")
    assert (two / "SV.matrix").read_bytes() == open(os.path.join(golden_dir, "sv_rand_b.SV.matrix"), "rb").read()
    for name in ("SV.matrix", "SV.dist", "SV.treefile"):
        assert (one / name).read_bytes() == (two / name).read_bytes(), name


def test_a_bad_record_in_one_workers_range_fails_the_whole_run(engine, tmp_path, capsys):
    text, _ = vcfgen.random_snp_vcf(3, 9, 400, multi_p=0.0, star_p=0.0)
    lines = text.rstrip("\n").split("\n")
    lines[-20] = "\t".join(lines[-20].split("\t")[:9] + ["7/7"] * 9)       # allele index out of range, in the last range
    src = tmp_path / "bad.vcf"
    src.write_text("\n".join(lines) + "\n")
    cfg = tmp_path / "run.cfg"
    for workers in ("threads", "processes"):
        _cfg(cfg, "snp_opt", src, tmp_path / "out", "treebest", f"gpus = 2\ngpu_workers = {workers}\n")
        with pytest.raises(SystemExit) as ei:
            run.main(["-s", str(cfg)])
        assert ei.value.code == 1
        out = capsys.readouterr().out
        if workers == "threads":                   # (worker processes print to the terminal, not through capsys)
            assert "allele index out of range" in out and "worker 1" in out


def test_more_workers_than_lines(engine, tmp_path):
    text, _ = vcfgen.random_snp_vcf(5, 12, 3, multi_p=0.0, star_p=0.0, indel_p=0.0, mono_p=0.0, blank_p=0.0)
    src = tmp_path / "tiny.vcf"
    src.write_text(text)
    one = _run(tmp_path, "1", "snp_opt", "-s", src, "treebest", "")
    four = _run(tmp_path, "4", "snp_opt", "-s", src, "treebest", "gpus = 4\n")
    for name in ("all_snp.fa", "SNP.dist", "SNP_treebest.NHX"):
        assert (one / name).read_bytes() == (four / name).read_bytes(), name


def test_worker_threads_leave_the_caller_single(engine, tmp_path):
    """After a `gpus = N` run on worker threads the calling thread is an ordinary one-GPU caller again, and the
    engine the test session holds still works on its device."""
    from phyloforge_b200 import multigpu
    text, _ = vcfgen.random_snp_vcf(23, 200, 900, miss=0.02, multi_p=0.0, star_p=0.0)
    src = tmp_path / "in.vcf"
    src.write_text(text)
    a = _run(tmp_path, "a", "snp_opt", "-s", src, "treebest", "gpus = 2\n")
    assert multigpu.worker_env() == (0, 1) and multigpu.comm_rank_world() == (0, 1)
    b = _run(tmp_path, "b", "snp_opt", "-s", src, "treebest", "")
    for name in ("all_snp.fa", "SNP.dist", "SNP_treebest.NHX"):
        assert (a / name).read_bytes() == (b / name).read_bytes(), name
    assert torch.cuda.current_device() == engine.device.index


@pytest.mark.parametrize("workers", ["threads", "processes"])
def test_bgzip_input_cut_over_workers_equals_one_gpu(engine, tmp_path, workers, capsys):
    """A bgzip file is cut at block starts and its lines dealt out along the cuts (vcf._BgzfLines): same artefacts as one
    GPU on the plain text; a gzip file bgzip did not write cannot be cut and says so."""
    import gzip
    text, _ = vcfgen.random_snp_vcf(29, 33, 1800, miss=0.03, multi_p=0.1, star_p=0.03)
    plain = tmp_path / "in.vcf"
    plain.write_text(text)
    bz = tmp_path / "in.vcf.gz"
    vcfgen.write_bgzf(bz, text.encode(), block=2500, seed=4)
    extra = "distance = ibs\nbootstrap = 8\n"
    one = _run(tmp_path, "1", "snp_opt", "-s", plain, "treebest", extra)
    many = _run(tmp_path, "3", "snp_opt", "-s", bz, "treebest", extra + f"gpus = 3\ngpu_workers = {workers}\n")
    for name in ("all_snp.fa", "SNP.dist", "SNP_treebest.NHX"):
        assert (one / name).read_bytes() == (many / name).read_bytes(), name
    rep1 = json.loads((one / "SNP.phyloforge_b200.json").read_text())
    repn = json.loads((many / "SNP.phyloforge_b200.json").read_text())
    for key in ("samples", "sites", "lines", "alignment_columns", "sites_not_2bit"):
        assert rep1[key] == repn[key], key
    if workers == "threads":
        gz = tmp_path / "plain.vcf.gz"
        with gzip.open(gz, "wb") as f:
            f.write(text.encode())
        cfg = tmp_path / "run_gz.cfg"
        _cfg(cfg, "snp_opt", gz, tmp_path / "out_gz", "treebest", "gpus = 2\n")
        capsys.readouterr()
        with pytest.raises(SystemExit) as ei:
            run.main(["-s", str(cfg)])
        assert ei.value.code == 1 and "bgzip" in capsys.readouterr().out
