"""cta_group::2 kind::mxf4 probe: exactness of the 256 x n CTA-pair instruction and clocks per instruction."""
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
for n, k, vals in ((128, 64, (-1, 0, 1)), (240, 64, (-1, 0, 1)), (240, 1024, (-1, 0, 1)), (256, 256, (-1, 0, 1)),
                   (240, 128, (-6, -4, -3, -2, -1, 0, 1, 2, 3, 4, 6)), (224, 1024, (1,))):
    a = rng.choice(np.array(vals, dtype=np.int8), size=(256, k))
    b = rng.choice(np.array(vals, dtype=np.int8), size=(n, k))
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    want = a.astype(np.int64) @ b.astype(np.int64).T
    dc = torch.full((256, n), -7.0, dtype=torch.float32, device="cuda")
    _lib.check_probe(_lib.load_probes().pf_selftest_umma_fp4_2cta(da.data_ptr(), db.data_ptr(), dc.data_ptr(), n, k, st))
    torch.cuda.synchronize()
    got = dc.cpu().numpy()
    ok = np.array_equal(got.astype(np.int64), want) and np.array_equal(got, np.rint(got))
    print(f"mxf4 2-CTA selftest 256x{n}x{k} values={vals[:3]}..: exact={ok} max|diff|={np.abs(got - want).max()}", flush=True)
    if not ok:
        bad = np.argwhere(got != want)
        print("  mismatches:", len(bad), "rows", np.unique(bad[:, 0])[:8], "cols", np.unique(bad[:, 1])[:8])
for n, bufs in ((128, 8), (192, 8), (224, 8), (240, 8), (240, 2)):
    v = C.c_double()
    pairs = C.c_int()
    _lib.check_probe(_lib.load_probes().pf_probe_umma_fp4_2cta_rate(n, 4000, bufs, C.byref(v), C.byref(pairs), st))
    print(f"mxf4 cta_group::2 256x{n}x64, 2 accumulators, {bufs} B tiles, {pairs.value} CTA pairs: {v.value:.1f} clocks per MMA", flush=True)
