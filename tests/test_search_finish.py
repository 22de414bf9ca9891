"""CPU: the host-side final ranking of libmnemo_cuda (sparse replay of glibc's merge sort with the
reference's non-transitive comparator, search.c:65-107,170-193) against the oracle, which calls libc
qsort on the dense per-entry array exactly as the reference does.  No GPU needed: mnemo_search_finish
is pure host code."""
import numpy as np

from mnemophonix_b200 import capi


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
