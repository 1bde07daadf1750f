"""Several GPUs behind `phyloforge -s / -S` (config key ``gpus = N``).

The reference shards the site axis of ONE VCF over worker processes inside ``snp_tree`` itself: ``cutFile`` cuts
the body into ``thread`` contiguous line ranges, a ``multiprocessing.Pool`` converts them, ``merge_seq``
concatenates the per-chunk sequences in chunk order (treetool/vcftree.py:67-94, 216-233). This module is the same
axis on GPUs: the entry point starts one worker per GPU; worker r converts the lines of byte range r of
the file (cut at line starts, ``vcf.byte_range_of_rank``) on its GPU, packs and counts its own columns, the int32
count matrices are summed with one allreduce (NCCL over NVLink), and rank 0 writes the artefacts: the workers'
character shards concatenated per sample in rank order, as ``merge_seq`` does with the chunk files (worker processes
exchange them through chunk FASTAs in ``{out}/working_dir`` like the reference's, worker threads in memory).

So that word ranges of the packed matrix (bootstrap blocks) mean the same sites for any number of GPUs, every
worker moves its packed columns onto the global 32-site word grid once the column counts of the ranks before it
are known (``pf_planes_shift``).

Two worker models behind the same code (every worker runs ``snp_tree`` / ``sv_tree`` itself and meets the others
in ``barrier`` / ``agree_ok`` / ``gather_ints`` / ``allreduce_sum`` below):

* ``gpu_workers = threads`` (default): one THREAD per GPU inside this process (``run_threads``). Each thread has
  its own engine, CUDA stream and pinned staging buffers on its device (the C ABI works on the calling thread's
  current device; ctypes calls, file reads and kernel launches release the interpreter lock). The count matrices
  are summed by one in-process NCCL allreduce over all devices (``torch.cuda.nccl``; the communicators are created
  by a helper thread while the workers convert their byte ranges), or - when workers share a device, or with
  ``PHYLOFORGE_B200_THREAD_REDUCE=p2p`` - by peer copies over NVLink: device r sums slice r of every peer's matrix
  and every device then fetches the finished slices. Nothing is started but threads, so `gpus = 8` costs the eight
  CUDA contexts and nothing else (the process model below spends ~14 s starting eight interpreters and their
  communicator before the first byte is converted, profiles/r02_vcf_mg8.json).
* ``gpu_workers = processes``: separate processes of the same command line (``python -m phyloforge_b200.run -s
  cfg``) told their rank through the environment, joined in a torch.distributed process group (NCCL); the parent
  only starts them, relays rank 0's output and exit status, and stops the rest if one fails. When there are more
  workers than visible GPUs (testing on a one-GPU box) the workers share devices and the allreduce goes through
  gloo on host copies - correct, slow, and said so on stdout.
"""
from __future__ import annotations

import os
import socket
import subprocess
import sys
import threading
import time
import traceback

ENV_RANK, ENV_WORLD, ENV_PORT = "PHYLOFORGE_B200_RANK", "PHYLOFORGE_B200_WORLD", "PHYLOFORGE_B200_PORT"
ENV_TRACE = "PHYLOFORGE_B200_TRACE"
ENV_THREAD_REDUCE = "PHYLOFORGE_B200_THREAD_REDUCE"     # nccl (default) | p2p: how worker THREADS sum their counts


def trace(label: str):
    """With PHYLOFORGE_B200_TRACE=1: seconds since this process was created, per worker, on stderr."""
    if not os.environ.get(ENV_TRACE):
        return
    try:
        import psutil
        t0 = psutil.Process().create_time()
    except Exception:
        t0 = _T_IMPORT
    print(f"[trace rank {worker_env()[0]}] +{time.time() - t0:7.3f} s  {label}", file=sys.stderr, flush=True)


_T_IMPORT = time.time()


class WorkerAborted(RuntimeError):
    """Another worker thread of the group failed: this one leaves at its next meeting point."""


class _Meeting:
    """Barrier of the worker threads. Unlike threading.Barrier, a meeting that everybody reached stays reached: a
    worker that fails right behind it (and aborts the group) does not take that meeting away from peers that have not
    woken up yet - they go on, say what they have to say, and fail at the NEXT meeting."""

    def __init__(self, parties: int):
        self.parties, self.count, self.generation, self.broken = parties, 0, 0, False
        self.cv = threading.Condition()

    def wait(self):
        with self.cv:
            if self.broken:
                raise WorkerAborted()
            generation = self.generation
            self.count += 1
            if self.count == self.parties:
                self.count = 0
                self.generation += 1
                self.cv.notify_all()
                return
            while generation == self.generation and not self.broken:
                self.cv.wait()
            if generation == self.generation:
                raise WorkerAborted()

    def abort(self):
        with self.cv:
            self.broken = True
            self.cv.notify_all()


class ThreadGroup:
    """Meeting points of the worker threads of one `gpus = N` run (gpu_workers = threads)."""

    def __init__(self, world: int):
        self.world = world
        self._barrier = _Meeting(world)
        self.slots = [None] * world
        self.status = [None] * world
        self.reduce_mode = None           # "nccl" | "p2p", decided by rank 0 at the first allreduce
        self.reduce_note = ""
        self._warm = None                 # helper thread creating the NCCL communicators
        self._warm_error = None

    def barrier(self):
        self._barrier.wait()

    def abort(self):
        self._barrier.abort()

    def exchange(self, rank: int, value):
        """Every worker's value, in rank order, on every worker."""
        self.slots[rank] = value
        self.barrier()
        out = list(self.slots)
        self.barrier()                    # nobody overwrites a slot somebody still reads
        return out

    # -- in-process NCCL: communicators over devices 0..world-1, created while the workers convert
    def start_nccl_warmup(self):
        import torch
        want = os.environ.get(ENV_THREAD_REDUCE, "nccl").strip().lower()
        if want not in ("nccl", "p2p"):
            want = "nccl"
        if want == "p2p":
            self.reduce_note = f"{ENV_THREAD_REDUCE}=p2p"
            return
        if not torch.cuda.is_available() or torch.cuda.device_count() < self.world:
            self.reduce_note = "workers share devices"
            return

        def warm():
            try:
                import torch.cuda.nccl as nccl
                ts = [torch.zeros(32, dtype=torch.int32, device=torch.device("cuda", d)) for d in range(self.world)]
                nccl.all_reduce(ts)
                for d in range(self.world):
                    torch.cuda.synchronize(d)
                trace("in-process NCCL communicators ready")
            except BaseException as e:  # noqa: BLE001 - reported by rank 0, which then sums through peer copies
                self._warm_error = e

        self._warm = threading.Thread(target=warm, name="phyloforge-nccl-init", daemon=True)
        self._warm.start()

    def decide_reduce_mode(self):
        """Rank 0, once: NCCL when its communicators came up, else peer copies (and why)."""
        if self.reduce_mode is not None:
            return
        if self._warm is None:
            self.reduce_mode = "p2p"
            return
        self._warm.join()
        if self._warm_error is not None:
            self.reduce_mode = "p2p"
            self.reduce_note = f"in-process NCCL failed ({self._warm_error})"
            print(f"gpus = {self.world}: {self.reduce_note}; the count matrices are summed through peer copies")
        else:
            self.reduce_mode = "nccl"


_TLS = threading.local()


def _thread_ctx():
    return getattr(_TLS, "ctx", None)       # (rank, ThreadGroup) of a worker thread


def worker_env():
    """(rank, world) of this worker: a worker thread's (run_threads), else the environment's (spawn_workers),
    else (0, 1)."""
    ctx = _thread_ctx()
    if ctx is not None:
        return ctx[0], ctx[1].world
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


def _status_of_exit(e: SystemExit) -> int:
    if e.code is None:
        return 0
    return e.code if isinstance(e.code, int) else 1


def _thread_main(group: ThreadGroup, rank: int, work):
    _TLS.ctx = (rank, group)
    status = 1
    try:
        rc = work()
        status = rc if isinstance(rc, int) else 0
    except SystemExit as e:                       # the workflow's own exits (bad record, failed child, ...)
        status = _status_of_exit(e)
    except WorkerAborted:
        status = None                             # somebody else failed and said why: that worker's status counts
    except BaseException:  # noqa: BLE001 - a worker thread must not die silently: the others would wait for it
        traceback.print_exc()
        status = 1
    finally:
        group.status[rank] = status
        if status != 0:
            group.abort()                         # (aborting a broken barrier again is harmless)
        _TLS.ctx = None


def run_threads(work, world: int, will_reduce: bool = True) -> int:
    """Run `work()` as `world` worker threads of this process, one per GPU (rank and world reach the workflow
    through worker_env); returns the exit status: 0 only if every worker returned 0, else the first failure's.
    will_reduce: the workflow sums count matrices over the workers (the GPU tree); without it (an external tree tool
    on the merged alignment file) no NCCL communicator is created."""
    group = ThreadGroup(world)
    if will_reduce:
        group.start_nccl_warmup()
    threads = [threading.Thread(target=_thread_main, args=(group, r, work), name=f"phyloforge-worker-{r}")
               for r in range(world)]
    # the workers spend their time inside the library, torch and file reads (no interpreter lock held); what is
    # left are short hand-overs between many threads, and a waiting thread asks for the lock only every switch
    # interval (5 ms by default: longer than a chunk's kernels take)
    interval = sys.getswitchinterval()
    sys.setswitchinterval(min(interval, 2e-4))
    try:
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    finally:
        sys.setswitchinterval(interval)
        if group._warm is not None:
            group._warm.join()            # (a run that failed before its first allreduce: never leave while NCCL initialises)
    for st in group.status:                       # the first worker (in rank order) that failed by itself
        if st not in (0, None):
            return st
    return 0 if all(st == 0 for st in group.status) else 1


def init_worker(rank: int, world: int) -> int:
    """Bind this worker to its GPU (and, for worker processes, join their process group); returns the CUDA device
    index of this worker."""
    import torch
    import torch.distributed as dist
    n_dev = torch.cuda.device_count()
    if n_dev < 1:
        from ._lib import PhyloForgeError
        raise PhyloForgeError("no CUDA device: phyloforge_b200 has no CPU fallback")
    device = rank % n_dev
    torch.cuda.set_device(device)
    if _thread_ctx() is not None:                 # a worker thread: the current device is per thread, nothing to join
        if world > n_dev and rank == 0:
            print(f"gpus = {world} but {n_dev} GPU(s) visible: workers share devices (testing only)")
        trace("cuda device set")
        return device
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
    ctx = _thread_ctx()
    if ctx is not None:
        return all(ctx[1].exchange(ctx[0], bool(ok)))
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
    ctx = _thread_ctx()
    if ctx is not None:
        return [list(v) for v in ctx[1].exchange(ctx[0], [int(x) for x in values])]
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


def gather_local(obj):
    """Worker threads: every worker's `obj` in rank order (the objects themselves: the workers share an address
    space). Worker processes: None - they exchange through the files in working_dir, as the reference's do."""
    ctx = _thread_ctx()
    if ctx is None:
        return None
    return ctx[1].exchange(ctx[0], obj)


def barrier():
    ctx = _thread_ctx()
    if ctx is not None:
        ctx[1].barrier()
        return
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def comm_rank_world():
    """(rank, world) of the group this worker exchanges with: worker threads, a torch.distributed process group
    (worker processes, torchrun), or (0, 1)."""
    ctx = _thread_ctx()
    if ctx is not None:
        return ctx[0], ctx[1].world
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def _sync(t):
    if t.is_cuda:
        import torch
        torch.cuda.current_stream(t.device).synchronize()


def thread_allreduce_sum(t):
    """Sum `t` (int32 counts, same shape on every worker thread, each on its worker's device) over the worker
    threads, in place. NCCL (one in-process call over all devices) when the workers sit on different GPUs; else -
    or for a tensor with padded rows - peer copies: worker r sums slice r of every peer's tensor, then every worker
    fetches the finished slices (a reduce-scatter and an all-gather written as device-to-device copies, over
    NVLink / NVSwitch between GPUs). Integer sums: the result does not depend on the route."""
    import torch
    from .sharding import shard_words
    rank, group = _thread_ctx()
    world = group.world
    _sync(t)                                       # this worker's tensor is complete before anybody reads it
    peers = group.exchange(rank, t)
    if rank == 0:
        group.decide_reduce_mode()
    group.barrier()
    if group.reduce_mode == "nccl" and t.is_cuda and all(p.is_contiguous() for p in peers):
        if rank == 0:
            import torch.cuda.nccl as nccl
            nccl.all_reduce(list(peers))           # in place on every device
            for p in peers:
                torch.cuda.synchronize(p.device)
        group.barrier()
        return t
    views = [p.view(-1) if p.is_contiguous() else p for p in peers]      # padded rows: cut along the first axis
    mine = views[rank]
    n = mine.shape[0]
    b, e = shard_words(n, world, rank)
    if e > b:
        acc = mine[b:e]
        for p in range(world):
            if p != rank:
                acc += views[p][b:e].to(acc.device, non_blocking=True)
    _sync(t)
    group.barrier()                                # every slice is finished in its owner's tensor
    for p in range(world):
        pb, pe = shard_words(n, world, p)
        if p != rank and pe > pb:
            mine[pb:pe].copy_(views[p][pb:pe], non_blocking=True)
    _sync(t)
    group.barrier()                                # nobody's tensor changes while a peer still reads it
    return t


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
