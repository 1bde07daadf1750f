"""Several GPUs behind `phyloforge -s / -S` (config key ``gpus = N``).

The reference shards the site axis of ONE VCF over worker processes inside ``snp_tree`` itself: ``cutFile`` cuts
the body into ``thread`` contiguous line ranges, a ``multiprocessing.Pool`` converts them, ``merge_seq``
concatenates the per-chunk sequences in chunk order (treetool/vcftree.py:67-94, 216-233). This module is the same
axis on GPUs: the entry point starts one worker process per GPU; worker r converts the lines of byte range r of
the file (cut at line starts, ``vcf.byte_range_of_rank``) on its GPU, packs and counts its own columns, the int32
count matrices are summed with one allreduce (NCCL over NVLink), and rank 0 writes the artefacts. The chunk
FASTAs land in ``{out}/working_dir`` as the reference's do and are concatenated in rank order.

So that word ranges of the packed matrix (bootstrap blocks) mean the same sites for any number of GPUs, every
worker moves its packed columns onto the global 32-site word grid once the column counts of the ranks before it
are known (``pf_planes_shift``).

Workers are separate processes of the same command line (``python -m phyloforge_b200.run -s cfg``) told their
rank through the environment; the parent only starts them, relays rank 0's output and exit status, and stops the
rest if one fails. When there are more workers than visible GPUs (testing on a one-GPU box) the workers share
devices and the allreduce goes through gloo on host copies — correct, slow, and said so on stdout.
"""
from __future__ import annotations

import os
import socket
import subprocess
import sys
import time

ENV_RANK, ENV_WORLD, ENV_PORT = "PHYLOFORGE_B200_RANK", "PHYLOFORGE_B200_WORLD", "PHYLOFORGE_B200_PORT"
ENV_TRACE = "PHYLOFORGE_B200_TRACE"


def trace(label: str):
    """With PHYLOFORGE_B200_TRACE=1: seconds since this process was created, per worker, on stderr."""
    if not os.environ.get(ENV_TRACE):
        return
    try:
        import psutil
        t0 = psutil.Process().create_time()
    except Exception:
        t0 = _T_IMPORT
    print(f"[trace rank {os.environ.get(ENV_RANK, '0')}] +{time.time() - t0:7.3f} s  {label}", file=sys.stderr, flush=True)


_T_IMPORT = time.time()


def worker_env():
    """(rank, world) of this process: (0, 1) unless it was started by spawn_workers."""
    try:
        rank, world = int(os.environ.get(ENV_RANK, "0")), int(os.environ.get(ENV_WORLD, "1"))
    except ValueError:
        return 0, 1
    if world < 1 or not (0 <= rank < world):
        return 0, 1
    return rank, world


def _free_port() -> int:
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def spawn_workers(argv, world: int, poll_s: float = 0.05) -> int:
    """Run `python -m phyloforge_b200.run *argv` as `world` worker processes; returns the exit status (0 only if
    every worker returned 0). A worker that fails takes the others down (they would wait in a collective forever)."""
    port = _free_port()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    procs = []
    for r in range(world):
        env = dict(os.environ)
        env.update({ENV_RANK: str(r), ENV_WORLD: str(world), ENV_PORT: str(port)})
        env["PYTHONPATH"] = root + (os.pathsep + env["PYTHONPATH"] if env.get("PYTHONPATH") else "")
        procs.append(subprocess.Popen([sys.executable, "-m", "phyloforge_b200.run", *argv], env=env))
    status = 0
    try:
        live = set(range(world))
        while live:
            for r in sorted(live):
                rc = procs[r].poll()
                if rc is None:
                    continue
                live.discard(r)
                if rc != 0 and status == 0:
                    status = rc
            if status != 0 and live:
                time.sleep(1.0)                      # let the others print their own message, then stop them
                for r in live:
                    procs[r].terminate()
                for r in live:
                    try:
                        procs[r].wait(timeout=10)
                    except subprocess.TimeoutExpired:
                        procs[r].kill()
                break
            time.sleep(poll_s)
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    return status


def init_worker(rank: int, world: int) -> int:
    """Join the workers' process group; returns the CUDA device index of this worker."""
    import torch
    import torch.distributed as dist
    n_dev = torch.cuda.device_count()
    if n_dev < 1:
        from ._lib import PhyloForgeError
        raise PhyloForgeError("no CUDA device: phyloforge_b200 has no CPU fallback")
    device = rank % n_dev
    torch.cuda.set_device(device)
    if not dist.is_initialized():
        port = os.environ.get(ENV_PORT, "29541")
        shared = world > n_dev
        if shared and rank == 0:
            print(f"gpus = {world} but {n_dev} GPU(s) visible: workers share devices and the count matrices are "
                  "summed through gloo on host copies (testing only)")
        kw = {} if shared else {"device_id": torch.device("cuda", device)}
        trace("cuda device set, joining the process group")
        dist.init_process_group("gloo" if shared else "nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank,
                                world_size=world, **kw)
        trace("process group up")
    return device


def agree_ok(ok: bool) -> bool:
    """True iff every worker says ok (one small allreduce); lets all workers leave together when one of them
    met a record the reference cannot convert."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return ok
    flag = torch.tensor([0 if ok else 1], dtype=torch.int32)
    if dist.get_backend() == "nccl":
        flag = flag.cuda()
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    return int(flag.item()) == 0


def gather_ints(values):
    """values (list of ints of this worker) -> [world][len(values)] on every worker."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [list(values)]
    t = torch.tensor(list(values), dtype=torch.int64)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [[int(x) for x in o.tolist()] for o in out]


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def globalise(engine, packed, rank: int, world: int):
    """Give the worker's shard its place in the whole matrix: packed.col0 / total_cols (columns of the packed
    2-bit matrix before this worker's / in all workers), packed.extra0 / total_extra (the same for the columns
    that do not fit 2 bits), totals of lines and columns for the report; and move the planes onto the global
    word grid: afterwards word j of packed.planes is global word packed.word0 + j."""
    import torch
    from ._lib import check
    n_extra = 0 if packed.extra_chars is None else int(packed.extra_chars.shape[0])
    table = gather_ints([packed.n_sites, n_extra, packed.n_lines, packed.n_columns, packed.n_not_2bit])
    packed.col0 = sum(t[0] for t in table[:rank])
    packed.total_cols = sum(t[0] for t in table)
    packed.extra0 = sum(t[1] for t in table[:rank])
    packed.total_extra = sum(t[1] for t in table)
    packed.total_lines = sum(t[2] for t in table)
    packed.total_columns = sum(t[3] for t in table)
    packed.total_not_2bit = sum(t[4] for t in table)
    shift = packed.col0 % 32
    packed.word0 = packed.col0 // 32
    if shift and packed.n_sites:
        n_words_dst = (packed.n_sites + shift + 31) // 32
        dst = engine.alloc_planes(packed.n_samples, n_words_dst)
        check(engine.lib.pf_planes_shift(packed.planes.data_ptr(), packed.n_words, dst.data_ptr(), n_words_dst,
                                         packed.n_samples, packed.n_sites, shift,
                                         torch.cuda.current_stream().cuda_stream))
        packed.planes, packed.n_words = dst, n_words_dst
    packed.lead = shift if packed.n_sites else 0          # code-3 columns in front of the worker's first site
    packed.n_sites_local = packed.n_sites
    packed.n_sites = packed.n_sites + packed.lead         # sites held by packed.planes, leading padding included
    return packed


def merge_chunk_fastas(paths, out_file: str, fmt: str):
    """Concatenate the workers' chunk FASTAs per sample, in rank order (what merge_seq does with the reference's
    chunk files, treetool/vcftree.py:145-171), into `out_file` (+ '.fa' or '.phy'). Every chunk lists the samples
    in header order with one sequence line each, so this is a line-wise concatenation."""
    names, parts = None, []
    for p in paths:
        with open(p, "rb") as f:
            lines = f.read().split(b"\n")
        heads, seqs = lines[0:-1:2], lines[1::2]
        if names is None:
            names = heads
        parts.append(seqs[:len(names)])
    names = names or []
    if fmt == "phylip":
        length = sum(len(p[0]) for p in parts) if names else 0
        with open(out_file + ".phy", "wb") as f:
            f.write(f"{len(names)} {length}\n".encode())
            for i, h in enumerate(names):
                f.write(h[1:] + b"  ")
                for p in parts:
                    f.write(p[i])
                f.write(b"\n")
        return out_file + ".phy"
    with open(out_file + ".fa", "wb") as f:
        for i, h in enumerate(names):
            f.write(h + b"\n")
            for p in parts:
                f.write(p[i])
            f.write(b"\n")
    return out_file + ".fa"
