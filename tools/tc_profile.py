"""Role-level wait accounting of the tensor-core pair-count kernel (diagnostic)."""
import ctypes as C, sys, time
sys.path.insert(0, ".")
import torch
from phyloforge_b200.engine import Engine
eng = Engine(0)
n, m = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3000, 2_000_000)
codes = eng.synth_codes(n, 0, m, seed=42)
nw = eng.n_words(m)
planes = eng.alloc_planes(n, nw)
eng.pack_codes(codes, n, planes, nw)
del codes
counts = eng.alloc_counts(n)
eng.pair_counts(planes, n, nw, counts, variant=3)
torch.cuda.synchronize()
for debug in (0, 2, 4, 6):
  buf = (C.c_longlong * 16)()
  eng.lib.pf_pair_counts_tc_profile(1 | (debug << 4), None)
  t0 = time.perf_counter()
  eng.pair_counts(planes, n, nw, counts, variant=3)
  torch.cuda.synchronize()
  dt = time.perf_counter() - t0
  eng.lib.pf_pair_counts_tc_profile(0, buf)
  v = list(buf)
  steps = max(v[8], 1)
  print(f"debug={debug} (1: MMA does not wait, 2: expansion idle, 4: no MMA)  n={n} m={m} wall {dt*1e3:.1f} ms; steps {steps}")
  print(f"MMA thread:   total/step {v[1]/steps:8.0f}  waiting for operands {v[0]/steps:8.0f}")
  print(f"A-row warp:   total/step {v[4]/steps:8.0f}  wait raw {v[2]/steps:8.0f}  wait free stage {v[3]/steps:8.0f}")
  print(f"B-row warp:   total/step {v[11]/steps:8.0f}  wait raw {v[9]/steps:8.0f}  wait free stage {v[10]/steps:8.0f}")
  print(f"fence+wait per step: A-row warp {v[12]/steps:8.0f}  B-row warp {v[13]/steps:8.0f}")
  print(f"TMA producer: total/step {v[6]/steps:8.0f}  waiting for free raw stage {v[5]/steps:8.0f}")
