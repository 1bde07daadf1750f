"""ctypes binding of libphyloforge_b200.so (include/phyloforge_b200.h).

Thin by design: every call passes raw device pointers (``tensor.data_ptr()``) and sizes.
There is no fallback: if the library is missing or a symbol is absent, importing the
binding raises, and every non-zero status becomes a ``PhyloForgeError``.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

LIB_NAME = "libphyloforge_b200.so"
LIB_PATH = Path(__file__).resolve().parent / LIB_NAME
PROBES_LIB_PATH = Path(__file__).resolve().parent / "libphyloforge_b200_probes.so"


class PhyloForgeError(RuntimeError):
    """A C-ABI call returned a non-zero status (message from pf_last_error)."""


_vp = C.c_void_p
_i64 = C.c_int64
_i32 = C.c_int
_u64 = C.c_uint64
_u32 = C.c_uint32
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_i64p = C.POINTER(C.c_int64)

# name -> (restype, argtypes). Must list every symbol include/phyloforge_b200.h declares
# (tests/test_abi.py checks the two against each other).
PROTOTYPES = {
    "pf_last_error": (C.c_char_p, []),
    "pf_version": (C.c_char_p, []),
    "pf_tile": (_i32, []),
    "pf_device_info": (_i32, [_ip, _ip, _ip, _i64p, _ip]),
    "pf_planes_len": (_i64, [_i64, _i64]),
    "pf_synth_codes": (_i32, [_vp, _i64, _i64, _i64, _i64, _u64, _i32, _u32, _vp]),
    "pf_pack_codes_u8": (_i32, [_vp, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _vp]),
    "pf_pack_codes_2bit": (_i32, [_vp, _i64, _i64, _i64, _vp, _i64, _i64, _i32, _vp]),
    "pf_planes_shift": (_i32, [_vp, _i64, _vp, _i64, _i64, _i64, _i32, _vp]),
    "pf_unpack_codes_u8": (_i32, [_vp, _i64, _i64, _i64, _i64, _vp, _i64, _vp]),
    "pf_transpose_u8": (_i32, [_vp, _i64, _vp, _i64, _i64, _vp, _i64, _vp]),
    "pf_line_index_work_bytes": (_i64, [_i64]),
    "pf_line_index": (_i32, [_vp, _i64, _vp, _i64, _i64p, _vp, _vp]),
    "pf_vcf_snp_parse": (_i32, [_vp, _i64, _vp, _i64, _i64, _vp, _vp, _i64, _vp, _vp]),
    "pf_vcf_sv_parse": (_i32, [_vp, _i64, _vp, _i64, _i64, _vp, _vp, _i64, _vp, _vp]),
    "pf_chars_to_codes": (_i32, [_vp, _i64, _i64, _i64, _vp, _i64, _vp, _vp]),
    "pf_chars_state_codes": (_i32, [_vp, _i64, _vp, _i64, _i64, _i32, _vp, _i64, _vp]),
    "pf_chars_allele_codes": (_i32, [_vp, _i64, _vp, _i64, _i64, _i32, _vp, _i64, _vp]),
    "pf_counts_add_extra": (_i32, [_vp, _i64, _i32, _vp, _i64, _vp, _i64, _i64, _vp]),
    "pf_vcf_select_rows": (_i32, [_vp, _i64, _i32, _i32, _vp, _i64, _i64p, _vp]),
    "pf_pair_counts": (_i32, [_vp, _i64, _i64, _i64, _i64, _vp, _i64, _i32, _vp]),
    "pf_pair_counts_tc_work_bytes": (_i64, [_i64, _i64]),
    "pf_pair_counts_tc_was_general": (_i32, [_vp, _i64, _i64, _ip, _vp]),
    "pf_pair_counts_tc": (_i32, [_vp, _i64, _i64, _i64, _i64, _vp, _i64, _vp, _i64, _i32, _ip, _vp]),
    "pf_pair_counts_tc_profile": (_i32, [_i32, _vp]),
    "pf_normalise": (_i32, [_vp, _i64, _i64, _i32, _vp, _i64, _vp, _vp]),
    "pf_newick_from_merges": (_i64, [_vp, _vp, _i64, _vp, _vp, _vp, _i64]),
    "pf_write_distance_matrix": (_i32, [C.c_char_p, _vp, _i64, _vp, _i64, _i32, _i32]),
    "pf_bootstrap_support": (_i32, [_vp, _vp, _i64, _i64, _vp, _i32]),
    "pf_bootstrap_weights": (_i32, [_u64, _i64, _i32, _i32, _vp, _vp]),
    "pf_bootstrap_counts": (_i32, [_vp, _i32, _i64, _vp, _i32, _vp, _vp]),
    "pf_nj_work_bytes": (_i64, [_i64]),
    "pf_nj_profile": (_i32, [_i32, _vp]),
    "pf_nj_lazy_profile": (_i32, [_i32, _vp]),
    "pf_pair_counts_tc_set_impl": (_i32, [_i32, _i32]),
    "pf_pair_counts_tc_timing": (_i32, [_i32, _vp]),
    "pf_nj_batch": (_i32, [_vp, _i64, _i64, _i32, _i64, _vp, _vp, _vp, _vp]),
    "pf_nj": (_i32, [_vp, _i64, _i64, _vp, _vp, _vp, _vp]),
    "pf_launch_count": (_i64, []),
}

# libphyloforge_b200_probes.so (include/phyloforge_b200_probes.h): measurement and self-test entry points,
# not part of the drop-in boundary; nothing in the product calls them
PROBE_PROTOTYPES = {
    "pf_last_error": (C.c_char_p, []),
    "pf_microbench_int_pipe": (_i32, [_i32, _i32, _dp, _dp, _vp]),
    "pf_microbench_mma": (_i32, [_i32, _i32, _dp, _vp]),
    "pf_selftest_umma": (_i32, [_vp, _vp, _vp, _i32, _i32, _vp]),
    "pf_probe_umma_rate": (_i32, [_i32, _i32, _i32, _i32, _dp, _vp]),
    "pf_probe_latency": (_i32, [_i32, _i32, _i32, _dp, _vp]),
    "pf_selftest_umma_fp4": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "pf_probe_umma_fp4_rate": (_i32, [_i32, _i32, _i32, _i32, _i32, _i32, _dp, _vp]),
    "pf_selftest_umma_fp4_2cta": (_i32, [_vp, _vp, _vp, _i32, _i32, _vp]),
    "pf_probe_umma_fp4_2cta_rate": (_i32, [_i32, _i32, _i32, _dp, _vp, _vp]),
    "pf_probe_umma_fp4_2cta_peak": (_i32, [_i32, _i32, _i32, _dp, _dp, _vp]),
}

_lib = None


def load(path: os.PathLike | None = None) -> C.CDLL:
    """Load the shared library once and bind every prototype. Raises if anything is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = Path(path) if path is not None else LIB_PATH
    if not p.exists():
        raise PhyloForgeError(
            f"{p} not found: build it with `python -m phyloforge_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(str(p))
    for name, (res, args) in PROTOTYPES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise PhyloForgeError(f"{p} does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


_probes = None


def load_probes() -> C.CDLL:
    """The probes / self-test library (bench.py, tests, tools). Raises if it is missing."""
    global _probes
    if _probes is None:
        if not PROBES_LIB_PATH.exists():
            raise PhyloForgeError(f"{PROBES_LIB_PATH} not found: build it with `python -m phyloforge_b200.build`")
        lib = C.CDLL(str(PROBES_LIB_PATH))
        for name, (res, args) in PROBE_PROTOTYPES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError as e:
                raise PhyloForgeError(f"{PROBES_LIB_PATH} does not export {name}") from e
            fn.restype = res
            fn.argtypes = args
        _probes = lib
    return _probes


def check_probe(status: int) -> None:
    if status != 0:
        msg = load_probes().pf_last_error().decode("utf-8", "replace")
        raise PhyloForgeError(f"status {status}: {msg}")


def check(status: int) -> None:
    if status != 0:
        msg = load().pf_last_error().decode("utf-8", "replace")
        raise PhyloForgeError(f"status {status}: {msg}")


def ptr(t) -> int:
    """Device (or host) pointer of a torch tensor / None."""
    if t is None:
        return None
    return t.data_ptr()
