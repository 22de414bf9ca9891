"""CPU: the host-side final ranking of libmnemo_cuda (sparse replay of glibc's merge sort with the
reference's non-transitive comparator, search.c:65-107,170-193) against the oracle, which calls libc
qsort on the dense per-entry array exactly as the reference does.  No GPU needed: mnemo_search_finish
is pure host code."""
import os
import sys

import numpy as np

from mnemophonix_b200 import capi

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ranking_ref import node_list as _node_list  # noqa: E402


def _finish_one(idx, score, n, n_entries):
    c = np.zeros(len(idx), capi.CAND_DTYPE)
    c["query"] = 0
    c["entry"] = idx
    c["score"] = score
    c["n_matches"] = n
    return capi.search_finish(c, None, 1, np.array([0, 1], np.uint32), n_entries)[0][0]


def test_selection_equals_libc_qsort_on_dense_array(oracle):
    rng = np.random.RandomState(0)
    for trial in range(2000):
        E = int(rng.choice([1, 2, 3, 7, 10, 11, 16, 37, 100, 257, 1000, 4099]))
        m = rng.randint(0, min(E, 25) + 1)
        idx = np.sort(rng.choice(E, m, replace=False))
        n = rng.randint(1, 30, m)
        avg = rng.choice([30, 31, 33, 34, 35, 36, 37, 38, 40, 44, 50], m) + rng.choice([0, 0.25, 0.5, 0.75], m)
        score = np.maximum(np.round(avg * n), 30 * n).astype(np.float32)
        dense_s = np.zeros(E, np.float32)
        dense_n = np.zeros(E, np.int32)
        dense_s[idx] = score
        dense_n[idx] = n
        assert _finish_one(idx, score, n, E) == oracle.select(dense_s, dense_n), (trial, E, idx, score, n)


def test_comparator_cycle_case(oracle):
    # A=(40,5), B=(37,10), C=(34,15) form a cycle under search.c:65-107
    for perm in ([0, 1, 2], [2, 0, 1], [1, 2, 0]):
        avg = np.array([40.0, 37.0, 34.0])[perm]
        n = np.array([5, 10, 15])[perm]
        E = 50
        idx = np.array([3, 17, 40])
        dense_s = np.zeros(E, np.float32)
        dense_n = np.zeros(E, np.int32)
        dense_s[idx] = avg * n
        dense_n[idx] = n
        assert _finish_one(idx, (avg * n).astype(np.float32), n, E) == oracle.select(dense_s, dense_n)


def test_more_than_ten_candidates_uses_the_window(oracle):
    rng = np.random.RandomState(4)
    for _ in range(300):
        E = 300
        m = 40
        idx = np.sort(rng.choice(E, m, replace=False))
        n = rng.randint(1, 14, m)
        avg = rng.uniform(30, 42, m)
        score = np.maximum(np.round(avg * n), 30 * n).astype(np.float32)
        dense_s = np.zeros(E, np.float32)
        dense_n = np.zeros(E, np.int32)
        dense_s[idx] = score
        dense_n[idx] = n
        assert _finish_one(idx, score, n, E) == oracle.select(dense_s, dense_n)


def test_last_group_merge_across_shards():
    # two shards, one query signature: shard 0's own last group must be scored after all because the
    # globally greatest (entry, sig) lives in shard 1 (search.c:147-165 is a global property)
    qstart = np.array([0, 1], np.uint32)
    cands = np.zeros(1, capi.CAND_DTYPE)
    cands[0] = (0, 1, 9 * 33.0, 9)                     # entry 1: 9 matches, average 33 (needs 10, or avg >= 35)
    last = np.zeros((2, 1), capi.LAST_DTYPE)
    last[0, 0] = (1, 1, 77, 3, 33)                     # shard 0 dropped entry 1 / sig 77: 3 hits, score 33
    last[1, 0] = (1, 5, 2, 2, 90)                      # shard 1 dropped entry 5 / sig 2 (the global last group)
    res = capi.search_finish(cands, last, 2, qstart, 8)
    assert res[0] == (1, 10 * 33.0, 10)                # the re-added match makes entry 1 qualify
    res = capi.search_finish(cands, None, 1, qstart, 8)
    assert res[0][0] == -7
    last[0, 0]["hits"] = 1                             # a single bucket hit is never deep-checked
    assert capi.search_finish(cands, last, 2, qstart, 8)[0][0] == -7


# ---- ranking tree: ten records per node are enough, and shards on node boundaries can be merged (host parts) ----

def test_node_lists_of_tree_aligned_shards_merge_to_the_reference_selection(oracle):
    rng = np.random.RandomState(21)
    for trial in range(400):
        E = int(rng.choice([1, 2, 3, 5, 8, 9, 16, 31, 100, 257, 1000]))
        m = rng.randint(0, E + 1) if trial % 3 else rng.randint(0, min(E, 12) + 1)
        idx = np.sort(rng.choice(E, m, replace=False))
        n = rng.randint(1, 30, m)
        avg = rng.choice([30, 31, 33, 34, 35, 36, 37, 38, 40, 44, 50], m) + rng.choice([0, 0.25, 0.5, 0.75], m)
        s = np.zeros(E, np.float32)
        c = np.zeros(E, np.int32)
        s[idx] = np.maximum(np.round(avg * n), 30 * n)
        c[idx] = n
        want = oracle.select(s, c)
        for shards in (1, 2, 3, 4, 8):
            cuts, node_lo = capi.tree_cuts(E, shards)
            n_nodes = len(node_lo) - 1
            assert cuts[0] == 0 and cuts[-1] == E and all(int(x) in set(node_lo.tolist()) for x in cuts)
            lists = np.zeros((n_nodes, 1, 10), capi.SEL_DTYPE)
            counts = np.zeros((n_nodes, 1), np.uint32)
            for i in range(n_nodes):
                lst = _node_list(s, c, int(node_lo[i]), int(node_lo[i + 1] - node_lo[i]))
                counts[i, 0] = len(lst)
                for k, (a, nn, e, sc) in enumerate(lst):
                    lists[i, 0, k] = (e, sc, nn, a)
            got = capi.select_merge(lists, counts)[0]
            assert got[0] == want, (trial, E, shards, got, want)


def test_tree_cuts_are_node_boundaries_of_the_merge_tree():
    """mnemo_tree_cuts: 2^ceil(log2 shards) nodes whose sizes come from msort's n/2 | n - n/2 split, contiguous groups of
    nodes per shard, every shard at least one node, everything covered, for awkward sizes too."""
    rng = np.random.RandomState(3)
    for trial in range(300):
        E = int(rng.choice([0, 1, 2, 3, 5, 7, 8, 9, 100, 1000, 10000, 12345, 1 << 20, rng.randint(1, 1 << 24)]))
        shards = int(rng.choice([1, 2, 3, 4, 5, 6, 7, 8, 16]))
        cuts, node_lo = capi.tree_cuts(E, shards)
        n_nodes = len(node_lo) - 1
        assert n_nodes >= shards and n_nodes & (n_nodes - 1) == 0 and n_nodes < 2 * max(shards, 1) + 1
        assert node_lo[0] == 0 and node_lo[-1] == E and np.all(np.diff(node_lo.astype(np.int64)) >= 0)
        sizes = np.diff(node_lo.astype(np.int64))
        assert sizes.max() - sizes.min() <= 1                      # floor / ceil of E / n_nodes
        # the split rule itself: every internal node's left child has floor(n / 2) entries
        lo, m = node_lo.astype(np.int64), n_nodes
        while m > 1:
            for i in range(0, n_nodes, m):
                n = lo[i + m] - lo[i]
                assert lo[i + m // 2] - lo[i] == n // 2
            m //= 2
        assert cuts[0] == 0 and cuts[-1] == E and len(cuts) == shards + 1
        assert all(int(c) in set(node_lo.tolist()) for c in cuts) and np.all(np.diff(cuts.astype(np.int64)) >= 0)
