"""Test helper: an independent Python replay of the ranking tree (search.c:65-107 comparator; glibc merge sort split
n/2 | n - n/2, left head taken when cmp <= 0; ten records per node).  Used by the CPU tests of the sharded protocols."""
import numpy as np


def cmp_scores(a, b):
    """search.c:65-107 on (avg float32, n)."""
    d = abs(int(np.float32(a[0]) - np.float32(b[0])))
    if d <= 3:
        if a[1] >= b[1] + 5:
            return -1
        if b[1] >= a[1] + 5:
            return 1
    if d < 0.5:
        if a[1] > b[1]:
            return -1
        if a[1] < b[1]:
            return 1
    if a[0] > b[0]:
        return -1
    if b[0] > a[0]:
        return 1
    return 0


def node_list(s, c, lo, n, top=10):
    """First `top` matched records of glibc's merge sort over positions lo .. lo+n-1 (split n/2 | n - n/2, left head
    taken when cmp <= 0), computed from the children's first `top` only."""
    if n <= 0:
        return []
    if n == 1:
        return [(np.float32(s[lo]) / np.float32(c[lo]), int(c[lo]), lo, float(s[lo]))] if c[lo] else []
    n1 = n // 2
    A, B = node_list(s, c, lo, n1, top), node_list(s, c, lo + n1, n - n1, top)
    out, ia, ib = [], 0, 0
    while len(out) < top and ia < len(A) and ib < len(B):
        if cmp_scores(A[ia], B[ib]) <= 0:
            out.append(A[ia]); ia += 1
        else:
            out.append(B[ib]); ib += 1
    while len(out) < top and ia < len(A):
        out.append(A[ia]); ia += 1
    while len(out) < top and ib < len(B):
        out.append(B[ib]); ib += 1
    return out
