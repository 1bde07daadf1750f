"""Probe legacy mma.sync rates on the box (b1 and.popc / xor.popc are software-emulated on sm_100a:
see profiles/r01_mma_probe.md)."""
import ctypes as C, json, sys
sys.path.insert(0, ".")
from phyloforge_b200 import _lib
lib = _lib.load_probes()
out = {}
for kind, name, iters in ((2, "mma.sync m16n8k32 s8 (IMMA.16832)", 4000), (0, "mma.sync m16n8k256 b1 and.popc (emulated)", 50),
                          (1, "mma.sync m16n8k256 b1 xor.popc (emulated)", 50)):
    v = C.c_double()
    rc = lib.pf_microbench_mma(kind, iters, C.byref(v), None)
    out[name] = v.value if rc == 0 else lib.pf_last_error().decode()
print(json.dumps(out, indent=1))
