"""PHYLIP square distance-matrix writer (the artefact `dnadist`/`treebest` users expect
beside the tree)."""
from __future__ import annotations

import os

import numpy as np


def write_distance_matrix(path: str, names, dist: np.ndarray, decimals: int = 10, threads: int = 0) -> None:
    """`n`, then per sample: label, two blanks, the n distances as "%.<decimals>f" - written by the library's host
    threads (csrc/artefacts.cu: pf_write_distance_matrix; 9 M numbers at 3,000 samples)."""
    import ctypes as C
    from . import _lib
    n = len(names)
    d = np.asarray(dist, dtype=np.float64)
    if d.ndim != 2 or d.shape[0] < n or d.shape[1] < n:
        raise ValueError(f"distance matrix of shape {d.shape} for {n} names")
    if not (d.strides[1] == 8 and d.strides[0] % 8 == 0 and d.strides[0] >= 8 * n):     # rows with a leading dimension
        d = np.ascontiguousarray(d)
    labels = (C.c_char_p * max(n, 1))(*[str(x).encode() for x in names])
    _lib.check(_lib.load().pf_write_distance_matrix(os.fsencode(path), labels, n, d.ctypes.data, d.strides[0] // 8,
                                                    decimals, threads))


def write_distance_matrix_py(path: str, names, dist: np.ndarray, fmt: str = "%.10f") -> None:
    """The same file with Python's own formatting: the statement of the format the tests hold the library to."""
    n = len(names)
    with open(path, "w") as f:
        f.write(f"{n}\n")
        for i in range(n):
            f.write(names[i] + "  " + " ".join(fmt % v for v in dist[i, :n]) + "\n")


def read_distance_matrix(path: str):
    with open(path) as f:
        n = int(f.readline().split()[0])
        names, rows = [], []
        for _ in range(n):
            parts = f.readline().split()
            names.append(parts[0])
            rows.append([float(x) for x in parts[1:1 + n]])
    return names, np.array(rows, dtype=np.float64)
