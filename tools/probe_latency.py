"""Latency of the NJ iteration's building blocks (pf_probe_latency)."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from phyloforge_b200._lib import check
from phyloforge_b200 import _lib
from phyloforge_b200.engine import Engine

eng = Engine(0)
st = torch.cuda.current_stream().cuda_stream
KINDS = ["L2 load chain", "store + fence.acq_rel.gpu", "__syncthreads (512 thr)", "tagged-word exchange",
         "cooperative grid.sync", "L2 load chain, remotely written lines", "batch of 16 L2 loads",
         "tcgen05.commit -> mbarrier wait", "cp.async.bulk 6 KB, issue to completion", "cp.async.bulk 6 KB, 7 in flight (per copy)",
         "mbarrier arrive -> try_wait wake-up (per hop)"]
only = [int(x) for x in sys.argv[1:]]
for kind, name in enumerate(KINDS):
    if only and kind not in only:
        continue
    for n_ctas in ((2, 32, 148) if kind in (3, 4) else (2,)):
        out = C.c_double(0)
        _lib.check_probe(_lib.load_probes().pf_probe_latency(kind, 2000, n_ctas, C.byref(out), st))
        print(f"{name:42s} ctas={n_ctas:4d}: {out.value:8.1f} ns/step", flush=True)
