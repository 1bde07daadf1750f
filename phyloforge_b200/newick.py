"""Host-side tree assembly: NJ merge list (pf_nj output) -> Newick text.

The reference obtains its SNP tree as the NHX/Newick text `treebest nj` prints
(treetool/script.py:359-360 -> ``{out}/SNP_treebest.NHX``); here the GPU kernel returns the
join order and branch lengths and this module serialises them (plain Newick is valid NHX).
"""
from __future__ import annotations

import re
from typing import Sequence

_SAFE = re.compile(r"^[^\s()\[\]':;,]+$")


def quote_label(name: str) -> str:
    """Newick label; quoted only when it holds a structural character."""
    if _SAFE.match(name):
        return name
    return "'" + name.replace("'", "''") + "'"


def merges_to_newick(merge, length, names: Sequence[str], fmt: str = "%.10g", support=None) -> str:
    """Serialise the pf_nj merge list as an unrooted Newick string with a trifurcating root.
    support[t] (optional, one value per row t < n-3) is written as the label of internal node n+t,
    i.e. of the edge above it (bootstrap support in percent).

    merge[t] = (a, b, 0) for t < n-3 (node n+t has children a, b with branch lengths
    length[t][0..1]); merge[n-3] = the three nodes of the final join, length[n-3] their
    branch lengths. Leaves are 0..n-1 and print as ``names[i]``.
    """
    n = len(names)
    if n == 1:
        return f"{quote_label(names[0])};"
    if n == 2:
        raise ValueError("two-leaf trees are handled by two_leaf_newick()")
    n_steps = n - 2
    if len(merge) != n_steps:
        raise ValueError(f"expected {n_steps} merge rows, got {len(merge)}")
    children: dict[int, tuple] = {}
    for t in range(n_steps - 1):
        a, b = int(merge[t][0]), int(merge[t][1])
        children[n + t] = ((a, float(length[t][0])), (b, float(length[t][1])))
    root = tuple((int(merge[n_steps - 1][k]), float(length[n_steps - 1][k])) for k in range(3))

    # iterative post-order so that 10^4-leaf caterpillar trees do not hit the recursion limit
    text: dict[int, str] = {}
    stack = [c for c, _ in root]
    while stack:
        v = stack[-1]
        if v < n:
            text[v] = quote_label(names[v])
            stack.pop()
            continue
        (a, la), (b, lb) = children[v]
        if a in text and b in text:
            label = "" if support is None else ("%g" % float(support[v - n]))
            text[v] = f"({text.pop(a)}:{fmt % la},{text.pop(b)}:{fmt % lb}){label}"
            stack.pop()
        else:
            if a not in text:
                stack.append(a)
            if b not in text:
                stack.append(b)
    parts = [f"{text[c]}:{fmt % l}" for c, l in root]
    return "(" + ",".join(parts) + ");"


def two_leaf_newick(names: Sequence[str], d: float, fmt: str = "%.10g") -> str:
    """Degenerate NJ on two samples: one edge of length d split evenly."""
    h = fmt % (d / 2.0)
    return f"({quote_label(names[0])}:{h},{quote_label(names[1])}:{h});"
