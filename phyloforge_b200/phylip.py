"""PHYLIP square distance-matrix writer (the artefact `dnadist`/`treebest` users expect
beside the tree)."""
from __future__ import annotations

import numpy as np


def write_distance_matrix(path: str, names, dist: np.ndarray, fmt: str = "%.10f") -> None:
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
