"""First-contact probe of a GPU box: host resources, GPU, and the integer-pipe microbenchmarks.

Run under gpurun:  python tools/probe_box.py > gpurun_out/probe.json
"""
import ctypes as C
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=60).stdout.strip()
    except Exception as e:  # noqa: BLE001
        return f"error: {e}"


def main():
    out = {
        "nproc": os.cpu_count(),
        "mem": sh("free -g | head -2"),
        "gpu": sh("nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.sm,power.limit --format=csv"),
        "cpu": sh("lscpu | grep -E 'Model name|Socket|Thread|Core'"),
        "ulimit_l": sh("ulimit -l"),
    }
    lib = C.CDLL(str(ROOT / "phyloforge_b200" / "libphyloforge_b200_probes.so"))
    lib.pf_microbench_int_pipe.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p]
    lib.pf_last_error.restype = C.c_char_p
    names = ["popc", "lop3", "add_s32(2 per IADD3)", "imad", "popc:lop3 1:6", "popc:imad 1:6", "lop3:imad 1:1",
             "mix_general(4lop3+3popc+3add)", "mix_nomiss(2lop3+2popc+2add)"]
    mb = {}
    for kind, name in enumerate(names):
        a, b = C.c_double(), C.c_double()
        rc = lib.pf_microbench_int_pipe(kind, 2000, C.byref(a), C.byref(b), None)
        if rc != 0:
            mb[name] = "error: " + lib.pf_last_error().decode()
        else:
            mb[name] = {"ops_per_s": a.value, "ops_per_clk_per_sm": b.value}
    out["microbench"] = mb
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
