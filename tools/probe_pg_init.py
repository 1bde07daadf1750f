"""Where do the seconds of a 2-worker start go? import torch / CUDA context / process group / first collectives."""
import os, sys, time
t00 = time.perf_counter()
import torch
import torch.distributed as dist
t_import = time.perf_counter() - t00
rank, world, port = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
backend = sys.argv[4] if len(sys.argv) > 4 else "nccl"
t0 = time.perf_counter(); torch.cuda.set_device(rank); torch.zeros(1, device="cuda"); torch.cuda.synchronize(); t_ctx = time.perf_counter() - t0
t0 = time.perf_counter()
kw = {"device_id": torch.device("cuda", rank)} if backend == "nccl" else {}
dist.init_process_group(backend, init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, **kw)
t_pg = time.perf_counter() - t0
t0 = time.perf_counter(); dist.barrier(); torch.cuda.synchronize(); t_bar = time.perf_counter() - t0
x = torch.ones((3, 3000, 3000), dtype=torch.int32, device="cuda") if backend == "nccl" else torch.ones((3, 3000, 3000), dtype=torch.int32)
t0 = time.perf_counter(); dist.all_reduce(x); torch.cuda.synchronize(); t_ar1 = time.perf_counter() - t0
t0 = time.perf_counter(); dist.all_reduce(x); torch.cuda.synchronize(); t_ar2 = time.perf_counter() - t0
t0 = time.perf_counter(); dist.barrier(); t_bar2 = time.perf_counter() - t0
print(f"{backend} rank {rank}: import {t_import:.2f} ctx {t_ctx:.2f} init_pg {t_pg:.2f} barrier1 {t_bar:.2f} allreduce1 {t_ar1:.3f} allreduce2 {t_ar2:.4f} barrier2 {t_bar2:.4f}", flush=True)
dist.destroy_process_group()
