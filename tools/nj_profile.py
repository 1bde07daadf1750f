"""Phase accounting of the NJ kernel (pf_nj_profile) for several matrix sizes."""
import ctypes as C
import sys
import time

import torch

sys.path.insert(0, ".")
from phyloforge_b200._lib import check
from phyloforge_b200.engine import Engine

eng = Engine(0)
NAMES = {0: ["scan", "cta-red+publish", "wait", "records+winner", "update"],
         2: ["scan+cta-red", "barrier1", "winner+update", "barrier2"]}
sizes = [int(x) for x in (sys.argv[1:] or ["1000", "3000", "6000", "10000"])]
for n in sizes:
    a = torch.rand((n, n), dtype=torch.float64, device=eng.device)
    d = (a + a.T) / 2
    d.fill_diagonal_(0)
    for general in (0, 2):
        check(eng.lib.pf_nj_profile(general, None))
        m0, l0 = eng.nj(d)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.nj(d)
        torch.cuda.synchronize()
        plain = time.perf_counter() - t0
        check(eng.lib.pf_nj_profile(general | 1, None))
        eng.nj(d)
        torch.cuda.synchronize()
        out = (C.c_ulonglong * 8)()
        check(eng.lib.pf_nj_profile(0, out))
        tot = sum(out[:5])
        print(f"n={n} {'general' if general else 'fast   '}: {plain*1e3:.1f} ms ({plain/n*1e6:.2f} us/iter); "
              f"instrumented {tot/1e6:.1f} ms: " +
              ", ".join(f"{nm} {out[i]/n/1e3:.2f}" for i, nm in enumerate(NAMES[general])) + " us/iter", flush=True)
        if general == 0:
            keep = (m0.clone(), l0.clone())
        else:
            print("   fast == general:", bool(torch.equal(keep[0], m0) and torch.equal(keep[1], l0)), flush=True)
