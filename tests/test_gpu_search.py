"""GPU: LSH build + batched search through the C ABI against the oracle / reference / golden results."""
import os

import numpy as np
import pytest

from mnemophonix_b200 import capi, synth

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(autouse=True, params=["buckets", "pairs", "pairs-cap5"])
def probe_path(request, monkeypatch):
    """Every test runs three times: candidates by streaming the probed buckets (k_search.cu), through the pair index
    (k_pairprobe.cu, forced on for these small databases), and through the pair index with room for five distinct
    candidates per round, so that nearly every query signature goes through the id ranges (narrowing of the runs,
    halving on overflow, widening after a sparse range) that only the heaviest ones need at full capacity."""
    monkeypatch.setenv("MNEMO_PAIR_INDEX", "0" if request.param == "buckets" else "1")
    if request.param == "pairs-cap5":
        monkeypatch.setenv("MNEMO_PAIR_CAP", "5")
    return request.param


@pytest.fixture(scope="module")
def corpus(ctx):
    pcms = [synth.synth_track(4000 + i, 20.0 + 3 * (i % 4), 1 + (i % 2)) for i in range(12)]
    sigs, st = ctx.index_pcm_batch(pcms)
    assert all(s == 0 for s in st)
    clips = [np.ascontiguousarray(p[6 * 44100 + 1234: 11 * 44100 + 1234]) for p in pcms]
    qs, st = ctx.index_pcm_batch(clips)
    return sigs, qs


def test_golden_search_results(ctx):
    g = np.load(os.path.join(HERE, "golden", "search_golden.npz"))
    n = int(g["n_tracks"][0])
    tracks = [g[f"track{i}"] for i in range(n)]
    db = capi.Db(ctx, tracks)
    queries = [g[f"query{i}"] for i in range(n + 1)] + [tracks[1]]
    res, (L, D) = db.search(queries)
    assert [r[0] for r in res] == g["results"].tolist()
    assert db.get_matches(g["gm0_sig"]) == [tuple(x) for x in g["gm0"].tolist()]
    assert db.get_matches(g["gm1_sig"]) == [tuple(x) for x in g["gm1"].tolist()]
    db.close()


def test_scores_and_stats_equal_oracle(ctx, oracle, corpus, probe_path):
    sigs, qs = corpus
    db = capi.Db(ctx, sigs)
    assert db.has_pair_index == probe_path.startswith("pairs")
    odb = oracle.make_db(sigs)
    assert db.table_size == sum(len(s) for s in sigs) // 2           # lsh.c:61
    res, (L, D) = db.search(qs)
    cands, last, qstart, _ = db.search_shard(qs)
    Lo = Do = 0
    for q, (query, r) in enumerate(zip(qs, res)):
        best, sc, l, d = oracle.search(odb, query)
        Lo += l
        Do += d
        assert r[0] == best == q                                     # identified track identical, and right
        mine = cands[cands["query"] == q]
        dense = np.zeros_like(sc)
        dense[mine["entry"], 0] = mine["score"]
        dense[mine["entry"], 1] = mine["n_matches"]
        assert np.array_equal(dense, sc)                             # per-entry (score, n_matches) identical
    assert (L, D) == (Lo, Do)                                        # probed nodes / deep checks identical
    db.close()


def test_self_search_reproduces_the_dropped_last_group(ctx, oracle, corpus):
    sigs, _ = corpus
    db = capi.Db(ctx, sigs)
    odb = oracle.make_db(sigs)
    res, _ = db.search([sigs[3], sigs[11]])
    for r, e in zip(res, (3, 11)):
        best, sc, _, _ = oracle.search(odb, sigs[e])
        assert r == (best, float(sc[e, 0]), int(sc[e, 1]))
        # a flush-corrected loop would count one more match per query signature whose own group is last
        assert sc[e, 1] < len(sigs[e]) + 10 * len(sigs[e])
    db.close()


def test_unrelated_and_empty_queries(ctx, corpus):
    sigs, _ = corpus
    db = capi.Db(ctx, sigs)
    rng = np.random.RandomState(1)
    noise = rng.randint(0, 255, (34, 100)).astype(np.uint8)
    res, _ = db.search([noise, np.zeros((0, 100), np.uint8), sigs[0][:5]])
    assert res[0][0] == -7 and res[1][0] == -7
    db.close()


def test_tiny_databases(ctx, oracle):
    one = [np.full((1, 100), 7, np.uint8)]
    db = capi.Db(ctx, one)                       # total_signatures/2 == 0: the reference would divide by zero
    assert db.table_size == 0
    assert db.search([one[0]])[0][0][0] == -7
    db.close()
    rng = np.random.RandomState(2)
    two = [rng.randint(0, 4, (3, 100)).astype(np.uint8), rng.randint(0, 4, (2, 100)).astype(np.uint8)]
    db = capi.Db(ctx, two)
    odb = oracle.make_db(two)
    res, _ = db.search([two[0], two[1]])
    assert [r[0] for r in res] == [oracle.search(odb, two[0])[0], oracle.search(odb, two[1])[0]]
    db.close()


def test_modulo_collisions_count_as_hits(ctx, oracle):
    # lsh.c:74,94: buckets are key % size; signatures whose keys differ by a multiple of size collide
    rng = np.random.RandomState(3)
    base = rng.randint(0, 200, (40, 100)).astype(np.uint8)
    entries = [base[:20].copy(), base[20:].copy()]
    size = 20
    q = base[5].copy()[None, :].repeat(12, 0)
    # alter bytes so that keys differ from entry 0 / sig 5 yet land in the same slots: add `size` to the
    # big-endian key of tables 0 and 1 (low byte + 20 stays < 256)
    q[:, 3] = (q[:, 3].astype(int) % 200 + size).astype(np.uint8)
    entries[0][5, 3] = q[0, 3] - size
    db = capi.Db(ctx, entries)
    odb = oracle.make_db(entries)
    assert db.table_size == size
    res, (L, D) = db.search([q])
    best, sc, Lo, Do = oracle.search(odb, q)
    assert (res[0][0], L, D) == (best, Lo, Do)
    db.close()


def test_sharded_search_equals_single_database(ctx, oracle, corpus):
    """Database split into contiguous entry ranges (one shard per GPU in production): every shard uses
    the GLOBAL table size, drops its own last group, and the merge restores all but the global one."""
    sigs, qs = corpus
    total = sum(len(s) for s in sigs)
    single = capi.Db(ctx, sigs)
    want, _ = single.search(qs + [sigs[2]])
    single.close()
    for cuts in ([0, 5, 12], [0, 3, 6, 9, 12], [0, 11, 12]):
        shards = [capi.Db(ctx, sigs[a:b], table_size=total // 2, entry_base=a) for a, b in zip(cuts[:-1], cuts[1:])]
        parts = [s.search_shard(qs + [sigs[2]]) for s in shards]
        cands = np.concatenate([p[0] for p in parts])
        last = np.stack([p[1] for p in parts])
        got = capi.search_finish(cands, last, len(shards), parts[0][2], len(sigs))
        assert got == want
        for s in shards:
            s.close()


@pytest.mark.parametrize("range_ids", [0, 1024, 128])
def test_long_buckets_take_the_partitioned_path(ctx, oracle, range_ids, monkeypatch):
    """Many near-identical signatures make one bucket thousands long and push the number of repeated candidates
    past the shared-memory short list: exercises the id-range mode of the probe kernel, with its list-free fallback
    (default range size: one pass with 6000 repeats), and with small ranges (test hook MNEMO_RANGE_IDS) the repeats
    collected over many passes and the walk over more than kMaxRangePasses ranges (128 ids per range: 51 ranges)."""
    if range_ids:
        monkeypatch.setenv("MNEMO_RANGE_IDS", str(range_ids))
    else:
        monkeypatch.delenv("MNEMO_RANGE_IDS", raising=False)
    rng = np.random.RandomState(5)
    proto = rng.randint(0, 250, 100).astype(np.uint8)
    e0 = np.repeat(proto[None, :], 6000, 0)
    e0[:, 8:] = rng.randint(0, 250, (6000, 92))          # tables 0,1 identical for all, the rest random
    e1 = rng.randint(0, 250, (500, 100)).astype(np.uint8)
    e1[:, :4] = proto[:4]                                 # only table 0 in common
    q = np.repeat(proto[None, :], 3, 0)
    q[1, 4:8] = 1
    db = capi.Db(ctx, [e0, e1])
    odb = oracle.make_db([e0, e1])
    res, (L, D) = db.search([q])
    best, sc, Lo, Do = oracle.search(odb, q)
    assert (res[0][0], L, D) == (best, Lo, Do)
    cands, _, _, _ = db.search_shard([q])
    dense = np.zeros_like(sc)
    dense[cands["entry"], 0] = cands["score"]
    dense[cands["entry"], 1] = cands["n_matches"]
    assert np.array_equal(dense, sc)
    db.close()


def test_dense_rows_equal_oracle_and_query_tiles(ctx, oracle, corpus, monkeypatch):
    """mnemo_search_scores: the dense (score, n_matches) rows of search.c:55-59 equal the oracle's for every query;
    with the accumulator budget shrunk to a few rows (MNEMO_ACC_BYTES) the batch is processed in several query tiles
    and nothing changes."""
    sigs, qs = corpus
    odb = oracle.make_db(sigs)
    want_rows = [oracle.search(odb, q)[1] for q in qs]
    for budget in (None, 8 * len(sigs) * 3):          # default: one tile; then tiles of three queries
        if budget:
            monkeypatch.setenv("MNEMO_ACC_BYTES", str(budget))
        else:
            monkeypatch.delenv("MNEMO_ACC_BYTES", raising=False)
        db = capi.Db(ctx, sigs)
        assert db.tile_queries == (3 if budget else (1 << 30) // (8 * len(sigs)))
        sc, nm = db.scores(qs)
        for q, rows in enumerate(want_rows):
            assert np.array_equal(sc[q], rows[:, 0].astype(np.uint32)) and np.array_equal(nm[q], rows[:, 1].astype(np.uint32))
        res, _ = db.search(qs)
        assert [r[0] for r in res] == list(range(len(qs)))
        cands, _, _, _ = db.search_shard(qs)
        assert sorted(set(cands["query"].tolist())) == list(range(len(qs)))
        db.close()


def test_get_matches_returns_every_member_of_long_chains(ctx, oracle):
    """A chain far longer than any fixed gather buffer comes back whole (lsh.c:89-112 never truncates)."""
    rng = np.random.RandomState(9)
    e0 = rng.randint(0, 250, (3000, 100)).astype(np.uint8)
    e0[:, :8] = 7                                        # tables 0 and 1 shared by all 3000
    db = capi.Db(ctx, [e0])
    got = db.get_matches(e0[5], cap=1 << 16)
    want = oracle.get_matches(oracle.make_db([e0]), e0[5])
    assert len(got) >= 6000 and got == want
    assert len(db.get_matches(e0[5], cap=16)) == 16      # a small caller buffer is filled, never overrun
    db.close()


# ------------------------------------------------------------------ final ranking on the GPU (k_select.cu) ----

def _select_cases(rng, n_cases, sizes, max_m, avgs=None):
    """Dense (score, n) rows with adversarial averages for search.c:65-107 (differences around 0.5, 3 and 5,
    match counts around the +5 rule and the 5/10 thresholds)."""
    rows = []
    for _ in range(n_cases):
        E = int(rng.choice(sizes))
        m = rng.randint(0, min(E, max_m) + 1)
        idx = np.sort(rng.choice(E, m, replace=False))
        n = rng.randint(1, 30, m)
        if avgs is None:
            avg = rng.choice([30, 31, 33, 34, 35, 36, 37, 38, 40, 44, 50], m) + rng.choice([0, 0.25, 0.5, 0.75], m)
        else:
            avg = rng.uniform(*avgs, m)
        s = np.zeros(E, np.uint32)
        c = np.zeros(E, np.uint32)
        s[idx] = np.maximum(np.round(avg * n), 30 * n).astype(np.uint32)
        c[idx] = n
        rows.append((s, c))
    return rows


@pytest.mark.gpu
@pytest.mark.parametrize("select_cap,global_tree", [(None, False), ("0", False), ("3", False), ("0", True)])
def test_gpu_ranking_equals_libc_qsort_on_dense_rows(ctx, oracle, monkeypatch, select_cap, global_tree):
    """The merge tree of k_select.cu against the oracle's qsort over the dense per-entry array, exactly as the
    reference sorts it (search.c:170): sparse, dense, more than ten candidates, sizes around powers of two.
    select_cap: rows with at most that many matched entries are ranked by the compacting kernel (default 2048), the
    others by the dense tree kernel — 0 sends everything to the dense kernel, 3 splits the small cases between the two;
    global_tree: the dense kernel for index ranges that do not fit shared memory instead of the shared-memory one."""
    if select_cap is not None:
        monkeypatch.setenv("MNEMO_SELECT_CAP", select_cap)
    if global_tree:
        monkeypatch.setenv("MNEMO_SELECT_GLOBAL", "1")
    rng = np.random.RandomState(11)
    cases = _select_cases(rng, 1500, [1, 2, 3, 4, 5, 7, 8, 9, 10, 11, 16, 17, 31, 33, 37, 100, 257, 1000, 1023, 1025, 4099], 25)
    cases += _select_cases(rng, 300, [300, 301, 511, 640], 40, avgs=(30, 42))
    cases += _select_cases(rng, 200, [64, 100, 999, 2048, 2500], 2500, avgs=(30, 45))        # most entries matched
    by_size = {}
    for s, c in cases:
        by_size.setdefault(len(s), []).append((s, c))
    for E, group in by_size.items():
        S = np.stack([g[0] for g in group])
        N = np.stack([g[1] for g in group])
        got = ctx.selftest_select(S, N)
        for k, (s, c) in enumerate(group):
            want = oracle.select(s.astype(np.float32), c.astype(np.int32))
            assert got[k][0] == want, (E, k, got[k], want)
            if want >= 0:
                assert got[k][1] == float(s[want]) and got[k][2] == int(c[want])


@pytest.mark.gpu
def test_gpu_ranking_comparator_cycle_and_entry_base(ctx, oracle):
    # A=(40,5), B=(37,10), C=(34,15) form a cycle under search.c:65-107
    for perm in ([0, 1, 2], [2, 0, 1], [1, 2, 0]):
        avg = np.array([40, 37, 34])[perm]
        n = np.array([5, 10, 15])[perm]
        s = np.zeros(50, np.uint32)
        c = np.zeros(50, np.uint32)
        s[[3, 17, 40]] = avg * n
        c[[3, 17, 40]] = n
        want = oracle.select(s.astype(np.float32), c.astype(np.int32))
        assert ctx.selftest_select(s[None], c[None])[0][0] == want
        # the same rows as the upper part of a larger index range (a shard searched on its own: the entries below
        # entry_base have no matches but still shape the merge tree)
        for base in (1, 7, 50, 333):
            full_s = np.concatenate([np.zeros(base, np.uint32), s])
            full_c = np.concatenate([np.zeros(base, np.uint32), c])
            want = oracle.select(full_s.astype(np.float32), full_c.astype(np.int32))
            assert ctx.selftest_select(s[None], c[None], entry_base=base)[0][0] == want


@pytest.mark.gpu
def test_ranked_shards_equal_single_database(ctx, oracle, corpus):
    """Sharded search ranked on the GPU (dist.RankedShardedSearcher's protocol, all shards in this process): shards on
    node boundaries of the ranking tree, one `has` byte per query signature exchanged, the last group of every shard
    but the highest one with candidates scored after all, ten records per query and node merged on the host."""
    sigs, qs = corpus
    queries = qs + [sigs[2]]
    total = sum(len(s) for s in sigs)
    single = capi.Db(ctx, sigs)
    want, _ = single.search(queries)
    single.close()
    odb = oracle.make_db(sigs)
    assert [w[0] for w in want] == [oracle.search(odb, q)[0] for q in queries]
    prepared = capi.Db.prepare(queries)
    for world in (1, 2, 3, 4, 8, 16):      # 16 shards for 12 entries: some shards and nodes are empty
        cuts, node_lo = capi.tree_cuts(len(sigs), world)
        n_nodes = len(node_lo) - 1
        shards = [capi.Db(ctx, sigs[int(cuts[r]):int(cuts[r + 1])], table_size=total // 2, entry_base=int(cuts[r]))
                  for r in range(world)]
        has = [s.dense_begin(prepared)[0] for s in shards]
        lists, counts = [], []
        for r, s in enumerate(shards):
            higher = np.zeros_like(has[r])
            for o in range(r + 1, world):
                higher |= has[o]
            n0, n1 = r * n_nodes // world, (r + 1) * n_nodes // world
            l, c = s.dense_lists(higher, node_lo[n0:n1 + 1])
            lists.append(l)
            counts.append(c)
        got = capi.select_merge(np.concatenate(lists), np.concatenate(counts))
        assert got == want, world
        for s in shards:
            s.close()


@pytest.mark.gpu
def test_device_resident_ranked_protocol_equals_single_database(ctx, corpus, monkeypatch):
    """The production form of the sharded search (mnemo_search_dense_begin_dev / _lists_dev / mnemo_select_merge_dev):
    flags, node lists and the merged answer never leave the device; here all shards live in this process and the two
    all-gathers are torch.cat.  Also through dist.RankedShardedSearcher with the queries cut into tiles."""
    import torch
    from mnemophonix_b200 import dist as mdist
    sigs, qs = corpus
    queries = qs + [sigs[2]]
    total = sum(len(s) for s in sigs)
    single = capi.Db(ctx, sigs)
    want, _ = single.search(queries)
    dense_want = single.scores(queries)
    single.close()
    prepared = capi.Db.prepare(queries)
    nq, nqs = len(queries), int(prepared[1][-1])
    dev = torch.device("cuda", 0)
    for world in (1, 2, 3, 5, 8, 16):
        cuts, node_lo = capi.tree_cuts(len(sigs), world)
        n_nodes = len(node_lo) - 1
        per = max((r + 1) * n_nodes // world - r * n_nodes // world for r in range(world))
        shards = [capi.Db(ctx, sigs[int(cuts[r]):int(cuts[r + 1])], table_size=total // 2, entry_base=int(cuts[r]))
                  for r in range(world)]
        has = [torch.zeros(nqs, dtype=torch.uint8, device=dev) for _ in range(world)]
        for r, s in enumerate(shards):
            s.dense_begin_dev(prepared, has[r].data_ptr())
        torch.cuda.synchronize()
        all_has = torch.cat(has)
        blocks = []
        for r, s in enumerate(shards):
            blk = torch.zeros(capi.shard_record_bytes(per, nq), dtype=torch.uint8, device=dev)
            n0, n1 = r * n_nodes // world, (r + 1) * n_nodes // world
            s.dense_lists_dev(all_has.data_ptr(), world, r, node_lo[n0:n1 + 1], blk.data_ptr(),
                              blk.data_ptr() + per * nq * 10 * capi.SEL_DTYPE.itemsize)
            blocks.append(blk)
        torch.cuda.synchronize()
        got = capi.select_merge_dev(ctx, torch.cat(blocks).data_ptr(), world, n_nodes, per, nq)
        assert got == want, world
        # the accumulators after the fix-up, shard by shard, are the single database's dense rows
        for q in (0, 5, nq - 1):
            row_s = np.concatenate([s.dense_read(q)[0] for s in shards])
            row_n = np.concatenate([s.dense_read(q)[1] for s in shards])
            assert np.array_equal(row_s, dense_want[0][q]) and np.array_equal(row_n, dense_want[1][q]), (world, q)
        for s in shards:
            s.close()
    # the searcher class on one rank (no collective), queries in tiles of four
    stream = torch.cuda.Stream()
    ctx2 = capi.Context(0, stream.cuda_stream)
    one = mdist.RankedShardedSearcher(ctx2, sigs, total // 2, len(sigs), 0, 1, device=dev, stream=stream, tile_queries=4)
    one.profile = True
    assert one.search(prepared) == want
    assert set(one.phase_ms) == set(one.PHASES) and one.phase_ms["probe"] > 0
    one.close()
    ctx2.close()


@pytest.mark.gpu
def test_page_locked_query_buffer_gives_the_same_answers(ctx, corpus):
    """Db.prepare(queries, ctx) puts the concatenated query signatures into page-locked memory (mnemo_host_alloc)."""
    sigs, qs = corpus
    db = capi.Db(ctx, sigs)
    want, stats = db.search(qs)
    pinned = capi.Db.prepare(qs, ctx)
    got, stats2 = db.search(pinned)
    assert got == want and stats2 == stats
    del pinned
    db.close()


def test_pair_index_is_built_lazily_and_changes_no_answer(ctx, corpus, monkeypatch):
    """mnemo_db_create only decides that the pair index pays (MNEMO_PAIR_INDEX=2: wanted, not forced); it is built when
    the database has answered MNEMO_PAIR_LAZY_AFTER query signatures — the streaming kernel answers until then — and the
    answers before, across and after the switch are the ones of a database that never has the index."""
    sigs, qs = corpus
    monkeypatch.setenv("MNEMO_PAIR_INDEX", "0")
    plain = capi.Db(ctx, sigs)
    want, stats = plain.search(qs)
    plain.close()
    n_first = sum(len(q) for q in qs[:2])
    monkeypatch.setenv("MNEMO_PAIR_INDEX", "2")
    monkeypatch.setenv("MNEMO_PAIR_LAZY_AFTER", str(n_first + 1))
    db = capi.Db(ctx, sigs)
    assert not db.has_pair_index
    got_a, _ = db.search(qs[:2])
    assert not db.has_pair_index                       # n_first query signatures so far: one short of the threshold
    got_b, _ = db.search(qs[2:])
    assert db.has_pair_index                           # built at the start of the batch that crosses it
    got_c, stats_c = db.search(qs)
    db.close()
    assert got_a + got_b == want and got_c == want and tuple(stats_c) == tuple(stats)
    monkeypatch.setenv("MNEMO_PAIR_LAZY_AFTER", "0")   # nothing to wait for: built at create
    db = capi.Db(ctx, sigs)
    assert db.has_pair_index
    db.close()
