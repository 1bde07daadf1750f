import torch, time
n = 7_520_000_000
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(393_216_000, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter()
    for o in range(0, n - d.numel(), d.numel()):
        d.copy_(h[o:o + d.numel()], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"H2D pinned: {(n // d.numel()) * d.numel() / dt / 1e9:.1f} GB/s, {dt*1e3:.1f} ms for {(n // d.numel()) * d.numel() / 1e9:.2f} GB")
x = torch.empty((3000, 3000), dtype=torch.float64, device="cuda")
t0 = time.perf_counter(); y = x.cpu(); dt = time.perf_counter() - t0
print(f"D2H 72 MB pageable .cpu(): {dt*1e3:.1f} ms")
p = torch.empty((3000, 3000), dtype=torch.float64).pin_memory()
t0 = time.perf_counter(); p.copy_(x, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"D2H 72 MB pinned: {dt*1e3:.1f} ms")
