"""kind::mxf4 tensor-core probe: exactness against integer dot products and clocks per instruction."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from phyloforge_b200._lib import check
from phyloforge_b200 import _lib
from phyloforge_b200.engine import Engine

eng = Engine(0)
st = torch.cuda.current_stream().cuda_stream
rng = np.random.default_rng(0)
for n, k, vals in ((128, 64, (-1, 0, 1)), (192, 256, (-1, 0, 1)), (128, 1024, (-1, 0, 1)), (192, 128, (-6, -4, -3, -2, -1, 0, 1, 2, 3, 4, 6)),
                   (192, 1024, (1,))):
    a = rng.choice(np.array(vals, dtype=np.int8), size=(128, k))
    b = rng.choice(np.array(vals, dtype=np.int8), size=(n, k))
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    want = a.astype(np.int64) @ b.astype(np.int64).T
    for ts in (0, 1):
        dc = torch.zeros((128, n), dtype=torch.float32, device="cuda")
        _lib.check_probe(_lib.load_probes().pf_selftest_umma_fp4(da.data_ptr(), db.data_ptr(), dc.data_ptr(), n, k, ts, st))
        torch.cuda.synchronize()
        got = dc.cpu().numpy()
        ok = np.array_equal(got.astype(np.int64), want) and np.array_equal(got, np.rint(got))
        print(f"mxf4 selftest n={n} k={k} A-in-{'TMEM' if ts else 'smem'} values={vals[:3]}..: exact={ok}  "
              f"max|diff|={np.abs(got - want).max()}", flush=True)
        if not ok:
            print(" got[0,:8]", got[0, :8], "want", want[0, :8])
for n, acc, ts, bufs, noise in ((128, 1, 0, 2, 0), (192, 2, 0, 8, 0), (192, 2, 1, 8, 0), (256, 1, 1, 8, 0),
                                (192, 2, 1, 8, 400), (192, 2, 1, 8, 200), (192, 2, 1, 8, 100), (192, 2, 1, 8, 50), (192, 2, 1, 8, 25),
                                (192, 2, 0, 8, 100), (192, 2, 0, 8, 50)):
    v = C.c_double()
    _lib.check_probe(_lib.load_probes().pf_probe_umma_fp4_rate(n, 4000, acc, ts, bufs, noise, C.byref(v), st))
    extra = f", LDS+STS noise {3 * 1024 / noise:.0f} B/clk" if noise else ""
    print(f"mxf4 128x{n}x64 A-in-{'TMEM' if ts else 'smem'}, {acc} accumulators, {bufs} B tiles{extra}: {v.value:.1f} clocks per MMA", flush=True)
