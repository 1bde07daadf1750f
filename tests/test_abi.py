"""The C-ABI library builds here (nvcc cross-compiles for sm_100a), loads, and exports every
symbol include/phyloforge_b200.h declares. No compute calls: there is no GPU in this container."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

from phyloforge_b200 import _lib, build

ROOT = Path(__file__).resolve().parents[1]
HEADER = ROOT / "include" / "phyloforge_b200.h"
PROBES_HEADER = ROOT / "include" / "phyloforge_b200_probes.h"


@pytest.fixture(scope="module")
def lib_path():
    return build.build()


def _declared(header=HEADER):
    text = header.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"PF_API\s+[\w\s\*]+?\b(pf_\w+)\s*\(", text))


def test_header_and_binding_agree():
    declared = _declared()
    assert declared, "no PF_API declarations found"
    assert declared == set(_lib.PROTOTYPES), (declared ^ set(_lib.PROTOTYPES))


def test_library_exports_every_declared_symbol(lib_path):
    out = subprocess.run(["nm", "-D", "--defined-only", str(lib_path)], capture_output=True, text=True, check=True)
    exported = {ln.split()[-1] for ln in out.stdout.splitlines() if " T " in ln}
    missing = _declared() - exported
    assert not missing, missing
    extra = {s for s in exported if s.startswith("pf_")} - _declared()
    assert not extra, f"exported but not declared in the header: {extra}"


def test_probes_are_a_separate_library(lib_path):
    """Microbenchmarks / self-tests live in libphyloforge_b200_probes.so with their own header: the product
    ABI exports no pf_probe_* / pf_selftest_* / pf_microbench_* symbol."""
    declared = _declared(PROBES_HEADER)
    assert declared == set(_lib.PROBE_PROTOTYPES), (declared ^ set(_lib.PROBE_PROTOTYPES))
    out = subprocess.run(["nm", "-D", "--defined-only", str(build.PROBES_LIB_PATH)], capture_output=True, text=True,
                         check=True)
    exported = {ln.split()[-1] for ln in out.stdout.splitlines() if " T " in ln and ln.split()[-1].startswith("pf_")}
    assert exported == declared, exported ^ declared
    assert not [s for s in _declared() if s.startswith(("pf_probe_", "pf_selftest_", "pf_microbench_"))]
    _lib.load_probes()


def test_library_loads_and_answers_without_a_gpu(lib_path):
    lib = _lib.load()
    assert b"sm_100a" in lib.pf_version()
    assert lib.pf_tile() == 64
    assert lib.pf_planes_len(65, 10) == 2 * 10 * 2 * 64
    assert lib.pf_nj_work_bytes(1000) > 12 * 1000
    assert lib.pf_line_index_work_bytes(1 << 20) >= 8 * 64
    assert lib.pf_launch_count() == 0


def test_sass_is_sm_100a_with_tma_bulk_copies(lib_path):
    out = subprocess.run(["cuobjdump", "-lelf", str(lib_path)], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_(\d+a?)", out.stdout))
    assert archs == {"100a"}, archs
    text = subprocess.run(["cuobjdump", "-sass", str(lib_path)], capture_output=True, text=True).stdout
    assert "UBLKCP" in text, "pair-count kernel should stage operands with TMA bulk copies"
    assert "POPC" in text


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(_lib.PhyloForgeError):
        _lib.load(tmp_path / "libphyloforge_b200.so")


def test_engine_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from phyloforge_b200.engine import Engine
    with pytest.raises(_lib.PhyloForgeError):
        Engine()
