import ctypes as C, sys
sys.path.insert(0, ".")
from phyloforge_b200 import _lib
lib = _lib.load_probes()
for n, ts, grp in ((256, 0, 1), (192, 0, 1), (192, 1, 1), (192, 1, 2), (192, 1, 8), (192, 1, 1000000), (128, 0, 1), (128, 1, 1), (128, 1, -3), (96, 1, -2), (96, 1, -4), (64, 1, -2), (64, 1, -4), (64, 1, -6)):
    v = C.c_double()
    rc = lib.pf_probe_umma_rate(n, 4000, ts, grp, C.byref(v), None)
    print(f"N={n} A-in-{'TMEM' if ts else 'smem'} group={grp}: {v.value:.1f} ticks per MMA (128x{n}x32)", "" if rc == 0 else lib.pf_last_error())
