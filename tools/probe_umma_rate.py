import ctypes as C, sys
sys.path.insert(0, ".")
from phyloforge_b200 import _lib
lib = _lib.load()
for n, ts in ((256, 0), (192, 0), (192, 1), (128, 0), (128, 1)):
    v = C.c_double()
    rc = lib.pf_probe_umma_rate(n, 4000, ts, C.byref(v), None)
    print(f"N={n} A-in-{'TMEM' if ts else 'smem'}: {v.value:.1f} ticks per MMA (128x{n}x32)", "" if rc == 0 else lib.pf_last_error())
