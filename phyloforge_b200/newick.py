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
    mg = merge.tolist() if hasattr(merge, "tolist") else [list(r) for r in merge]
    ln = length.tolist() if hasattr(length, "tolist") else [list(r) for r in length]
    sup = None if support is None else [("%g" % float(x)) for x in support]
    # one pass over a token stack (no recursion, no repeated copying of subtree strings: 10^4-leaf
    # caterpillar trees stay linear); a stack item is a finished piece of text or a node to expand
    out = []
    root = mg[n_steps - 1]
    rl = ln[n_steps - 1]
    stack = [");", int(root[2]), f":{fmt % rl[1]},", int(root[1]), f":{fmt % rl[0]},", int(root[0]), "("]
    while stack:
        x = stack.pop()
        if x.__class__ is str:
            out.append(x)
        elif x < n:
            out.append(quote_label(names[x]))
        else:
            t = x - n
            a, b = mg[t][0], mg[t][1]
            la, lb = ln[t][0], ln[t][1]
            stack.append(f":{fmt % lb})" if sup is None else f":{fmt % lb}){sup[t]}")
            stack.append(int(b))
            stack.append(f":{fmt % la},")
            stack.append(int(a))
            stack.append("(")
    text = "".join(out)
    # the root's third branch length sits before the closing parenthesis
    return text[:-2] + f":{fmt % rl[2]});"


def merges_to_newick_native(lib, merge, length, names: Sequence[str], support=None) -> str:
    """merges_to_newick(fmt="%.10g") through the library's host-side serialiser
    (pf_newick_from_merges): same text, a few times faster for thousands of leaves."""
    import ctypes as C

    import numpy as np
    n = len(names)
    if n < 3:
        return merges_to_newick(merge, length, names, support=support)
    mg = np.ascontiguousarray(merge, dtype=np.int32)
    ln = np.ascontiguousarray(length, dtype=np.float64)
    if mg.shape != (n - 2, 3) or ln.shape != (n - 2, 3):
        raise ValueError(f"expected {n - 2} merge rows")
    enc = [quote_label(x).encode() for x in names]
    labels = (C.c_char_p * n)(*enc)
    sup = None
    if support is not None:
        sup = np.ascontiguousarray(support, dtype=np.float64)
        if sup.shape[0] < n - 3:
            raise ValueError("one support value per internal edge expected")
    sup_p = sup.ctypes.data if sup is not None else None
    cap = 48 * n + sum(map(len, enc)) + 64
    buf = C.create_string_buffer(cap)
    need = lib.pf_newick_from_merges(mg.ctypes.data, ln.ctypes.data, n, labels, sup_p, buf, cap)
    if need >= cap:
        cap = need + 1
        buf = C.create_string_buffer(cap)
        need = lib.pf_newick_from_merges(mg.ctypes.data, ln.ctypes.data, n, labels, sup_p, buf, cap)
    if need < 0:
        raise ValueError(lib.pf_last_error().decode())
    return buf.raw[:need].decode()


def two_leaf_newick(names: Sequence[str], d: float, fmt: str = "%.10g") -> str:
    """Degenerate NJ on two samples: one edge of length d split evenly."""
    h = fmt % (d / 2.0)
    return f"({quote_label(names[0])}:{h},{quote_label(names[1])}:{h});"
