"""Restatement of NumPy's float64 pairwise summation (TEST INFRASTRUCTURE).

``np.add.reduce`` over a contiguous float64 vector recurses: fewer than 8
elements are added left to right; up to 128 elements are summed with eight
stride-8 partial sums combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) followed
by a sequential tail; longer ranges are split at ``n//2`` rounded down to a
multiple of 8.  The CUDA reductions reproduce this
tree, so this module is both the checker for that claim and the executable
description of the "leaf slot" decomposition the kernels use.
"""
from __future__ import annotations

import numpy as np

LEAF_MAX = 128


def leaf_sum(v: np.ndarray) -> float:
    n = len(v)
    if n < 8:
        s = np.float64(0.0)          # NumPy: res = 0.; res += a[i]
        for i in range(n):
            s = np.float64(s + v[i])
        return s
    r = [np.float64(v[k]) for k in range(8)]
    i = 8
    while i + 8 <= n:
        for k in range(8):
            r[k] = np.float64(r[k] + v[i + k])
        i += 8
    s = np.float64(np.float64(np.float64(r[0] + r[1]) + np.float64(r[2] + r[3]))
                   + np.float64(np.float64(r[4] + r[5]) + np.float64(r[6] + r[7])))
    while i < n:
        s = np.float64(s + v[i])
        i += 1
    return s


def pairwise_sum(v: np.ndarray) -> float:
    n = len(v)
    if n <= LEAF_MAX:
        return leaf_sum(v)
    n2 = n // 2
    n2 -= n2 % 8
    return np.float64(pairwise_sum(v[:n2]) + pairwise_sum(v[n2:]))


def tree_depth(n: int) -> int:
    """Depth D of the complete binary tree that covers the recursion for n."""
    d, s = 0, n
    while s > LEAF_MAX:
        h = s // 2
        h -= h % 8
        s = s - h          # the right child is the larger one
        d += 1
    return d


def leaf_slot(n: int, depth: int, slot: int):
    """(offset, size) of leaf slot ``slot`` in the depth-``depth`` complete tree.

    A range that stops splitting early keeps itself as its left descendant and
    an empty (size 0) right descendant, so every slot is reachable by walking
    the bits of ``slot`` from the most significant one.
    """
    off, s = 0, n
    for lvl in range(depth - 1, -1, -1):
        bit = (slot >> lvl) & 1
        if s > LEAF_MAX:
            h = s // 2
            h -= h % 8
            if bit:
                off += h
                s -= h
            else:
                s = h
        elif bit:
            return off, 0
    return off, s


def slot_tree_sum(v: np.ndarray) -> float:
    """Sum via the leaf-slot decomposition; must equal ``pairwise_sum`` bitwise."""
    n = len(v)
    d = tree_depth(n)
    vals, ok = [], []
    for slot in range(1 << d):
        off, s = leaf_slot(n, d, slot)
        ok.append(s > 0)
        vals.append(leaf_sum(v[off:off + s]) if s > 0 else np.float64(0.0))
    while len(vals) > 1:
        nv, nk = [], []
        for i in range(0, len(vals), 2):
            if ok[i] and ok[i + 1]:
                nv.append(np.float64(vals[i] + vals[i + 1])); nk.append(True)
            elif ok[i]:
                nv.append(vals[i]); nk.append(True)
            else:
                nv.append(np.float64(0.0)); nk.append(False)
        vals, ok = nv, nk
    return vals[0]
