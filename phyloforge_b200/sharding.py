"""Site sharding across GPUs: the site axis is cut into contiguous,
word-aligned (32-site) ranges — the axis the reference chunks across worker processes
(treetool/vcftree.py:77-93). Each rank counts its own sites for all sample pairs; the
integer count matrices are summed with one allreduce."""
from __future__ import annotations


def shard_words(n_words: int, world: int, rank: int):
    """Contiguous word range [begin, end) of `rank`; the first n_words % world ranks get one extra."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} of {world}")
    base, extra = divmod(n_words, world)
    begin = rank * base + min(rank, extra)
    end = begin + base + (1 if rank < extra else 0)
    return begin, end


def shard_sites(n_sites: int, world: int, rank: int):
    """Site range [begin, end) of `rank`, cut at 32-site word boundaries."""
    n_words = (n_sites + 31) // 32
    wb, we = shard_words(n_words, world, rank)
    return min(wb * 32, n_sites), min(we * 32, n_sites)


def allreduce_counts(counts):
    """Sum the site shards' count matrices over all ranks, in place (NCCL over NVLink on the
    GPUs; gloo in the CPU tests). Integer addition is associative, so the totals are
    bit-identical for any number of ranks. Worker threads of one process (gpus = N behind the entry point,
    multigpu.run_threads) sum through multigpu.thread_allreduce_sum. No-op without a group."""
    from . import multigpu
    if multigpu._thread_ctx() is not None:
        return multigpu.thread_allreduce_sum(counts)
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if counts.is_cuda and dist.get_backend() != "nccl":
            # workers that share a GPU (more workers than devices: testing) sum through host copies
            host = counts.cpu()
            dist.all_reduce(host, op=dist.ReduceOp.SUM)
            counts.copy_(host)
        else:
            dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return counts
