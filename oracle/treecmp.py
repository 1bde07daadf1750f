"""TEST INFRASTRUCTURE ONLY — Newick parser and tree comparators:
Robinson-Foulds distance on unrooted bipartition sets and the maximum absolute branch-length
difference matched by bipartition. See oracle/__init__.py."""
from __future__ import annotations


class Node:
    __slots__ = ("name", "length", "children")

    def __init__(self):
        self.name = None
        self.length = None
        self.children = []


def parse_newick(text: str) -> Node:
    """Iterative Newick parser (labels may be single-quoted; NHX comments are skipped)."""
    s = text.strip()
    if not s.endswith(";"):
        raise ValueError("Newick string must end with ';'")
    root = Node()
    cur = root
    stack = []
    i, n = 0, len(s)
    while i < n:
        ch = s[i]
        if ch == "(":
            child = Node()
            cur.children.append(child)
            stack.append(cur)
            cur = child
            i += 1
        elif ch == ",":
            cur = stack[-1]
            child = Node()
            cur.children.append(child)
            cur = child
            i += 1
        elif ch == ")":
            cur = stack.pop()
            i += 1
        elif ch == ":":
            j = i + 1
            while j < n and s[j] not in ",()[;":
                j += 1
            cur.length = float(s[i + 1:j])
            i = j
        elif ch == "[":
            i = s.index("]", i) + 1
        elif ch == ";":
            break
        elif ch == "'":
            j = i + 1
            out = []
            while True:
                if s[j] == "'":
                    if j + 1 < n and s[j + 1] == "'":
                        out.append("'")
                        j += 2
                        continue
                    break
                out.append(s[j])
                j += 1
            cur.name = "".join(out)
            i = j + 1
        elif ch.isspace():
            i += 1
        else:
            j = i
            while j < n and s[j] not in ":,()[;":
                j += 1
            cur.name = s[i:j].strip()
            i = j
    if stack:
        raise ValueError("unbalanced parentheses")
    return root


def leaves(root: Node):
    out, st = [], [root]
    while st:
        v = st.pop()
        if v.children:
            st.extend(v.children)
        else:
            out.append(v.name)
    return out


def splits(root: Node):
    """{canonical bipartition (frozenset of leaf names on the side without the reference
    leaf): branch length}. Trivial (single-leaf) splits are included so every edge is
    compared; a bifurcating root's two edges are merged into one split."""
    all_leaves = sorted(leaves(root))
    if len(set(all_leaves)) != len(all_leaves):
        raise ValueError("duplicate leaf names")
    ref = all_leaves[0]
    full = frozenset(all_leaves)
    below = {}
    order, st = [], [root]
    while st:
        v = st.pop()
        order.append(v)
        st.extend(v.children)
    for v in reversed(order):
        if v.children:
            acc = set()
            for c in v.children:
                acc |= below[id(c)]
            below[id(v)] = frozenset(acc)
        else:
            below[id(v)] = frozenset([v.name])
    out = {}
    for v in order:
        if v is root:
            continue
        side = below[id(v)]
        key = side if ref not in side else full - side
        if not key:
            continue
        out[key] = out.get(key, 0.0) + (v.length if v.length is not None else 0.0)
    return out


def robinson_foulds(a: Node, b: Node) -> int:
    sa = {k for k in splits(a) if len(k) > 1}
    sb = {k for k in splits(b) if len(k) > 1}
    if set(leaves(a)) != set(leaves(b)):
        raise ValueError("trees have different leaf sets")
    return len(sa ^ sb)


def max_branch_diff(a: Node, b: Node) -> float:
    """Max |length difference| over the splits the two trees share (all, when RF = 0)."""
    sa, sb = splits(a), splits(b)
    common = set(sa) & set(sb)
    if not common:
        return float("inf")
    return max(abs(sa[k] - sb[k]) for k in common)


def compare(newick_a: str, newick_b: str):
    ta, tb = parse_newick(newick_a), parse_newick(newick_b)
    return robinson_foulds(ta, tb), max_branch_diff(ta, tb)
