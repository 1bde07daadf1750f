#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "nj or end_to_end or smoke" 2>&1 | tail -4
timeout 300 python - <<'PY'
import sys, time, torch, numpy as np
sys.path.insert(0, ".")
from phyloforge_b200.engine import Engine
eng = Engine(0)
for n in (500, 1000, 2000, 3000, 5000, 10000):
    rng = np.random.default_rng(n)
    a = torch.rand((n, n), dtype=torch.float64, device=eng.device)
    d = (a + a.T) / 2
    d.fill_diagonal_(0)
    eng.nj(d)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        eng.nj(d)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(f"NJ n={n}: {dt*1e3:.1f} ms (incl. matrix clone), {dt/n*1e6:.2f} us/iteration")
PY
