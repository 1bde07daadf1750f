"""Build libphyloforge_b200.so (the product: the path's kernels + the C ABI of include/phyloforge_b200.h)
and libphyloforge_b200_probes.so (microbenchmarks / instruction self-tests, include/phyloforge_b200_probes.h;
used by bench.py's roofline denominators, tests and tools, never by the product) in-tree with nvcc for sm_100a.

Called by ``__graft_entry__.build()`` and usable as ``python -m phyloforge_b200.build``.
The library is the product; there is no CPU fallback and no JIT cache: the built
``.so`` lives next to this file so it travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libphyloforge_b200.so"
PROBES_LIB_PATH = PKG_DIR / "libphyloforge_b200_probes.so"
PROBE_SOURCES = {"latency_probe.cu", "microbench.cu", "umma_selftest.cu", "umma_fp4_probe.cu", "umma_fp4_2cta_probe.cu"}
STAMP = PKG_DIR / "csrc" / ".build_stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-fvisibility=hidden",
    "--fmad=false",      # the fp64 NJ/normalise path must match the oracle bit for bit: no contraction
    "-Xptxas", "-v",
    "-rdc=false",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libphyloforge_b200.so cannot be built")
    return exe


def sources() -> list[Path]:
    """Sources of the product library."""
    return sorted(p for p in CSRC.glob("*.cu") if p.name not in PROBE_SOURCES)


def probe_sources() -> list[Path]:
    return sorted(p for p in CSRC.glob("*.cu") if p.name in PROBE_SOURCES)


def _digest() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) +
                    [PKG_DIR.parent / "include" / "phyloforge_b200.h",
                     PKG_DIR.parent / "include" / "phyloforge_b200_probes.h"]):
        if p.exists():
            h.update(p.name.encode())
            h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu under csrc/ into one shared library. Returns its path."""
    digest = _digest()
    if (not force and LIB_PATH.exists() and PROBES_LIB_PATH.exists() and STAMP.exists()
            and STAMP.read_text().strip() == digest):
        return LIB_PATH
    objdir = PKG_DIR / "csrc" / "build"
    objdir.mkdir(exist_ok=True)
    nvcc = _nvcc()
    inc = ["-I", str(CSRC), "-I", str(PKG_DIR.parent / "include")]
    procs = []
    objs, probe_objs = [], []
    jobs = [(src, objdir / (src.stem + ".o"), [], objs) for src in sources()]
    jobs += [(src, objdir / (src.stem + ".o"), [], probe_objs) for src in probe_sources()]
    # the probes library carries its own copy of the error plumbing (api.cu without the product's exports)
    jobs.append((CSRC / "api.cu", objdir / "api_probes.o", ["-DPF_PROBES_LIB"], probe_objs))
    for src, obj, extra, into in jobs:
        into.append(str(obj))
        cmd = [nvcc, *NVCC_FLAGS, *extra, *inc, "-c", str(src), "-o", str(obj)]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    logs = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        logs.append(f"==> {src.name}\n{out}")
        if p.returncode != 0:
            failed = True
    (objdir / "ptxas.log").write_text("\n".join(logs))
    if failed or verbose:
        sys.stderr.write("\n".join(logs) + "\n")
    if failed:
        raise RuntimeError("nvcc failed; see log above")
    for lib, lib_objs in ((LIB_PATH, objs), (PROBES_LIB_PATH, probe_objs)):
        subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(lib), *lib_objs],
                       check=True)
    STAMP.write_text(digest)
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
