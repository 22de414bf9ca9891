"""CPU, world_size 2 over gloo: the multi-rank plumbing of mnemophonix_b200.dist.

Index sharding is pure bookkeeping (no collective on the data path).  For the sharded search the GPU
probe of each rank is stood in for by the ORACLE run on that rank's shard (same global table size, same
"drop my own last group" rule), so what is tested is exactly what runs across GPUs in production: the
all-gather of candidate / last-group records and the host-side exact merge + ranking."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_corpus():
    rng = np.random.RandomState(11)
    entries = []
    for e in range(9):
        n = int(rng.randint(30, 90))
        entries.append(rng.randint(0, 12, (n, 100)).astype(np.uint8))     # small alphabet: plenty of bucket hits
    queries = []
    for e in (1, 4, 8, 8):
        q = entries[e][5:30].copy()
        flip = rng.rand(*q.shape) < 0.3
        q[flip] = rng.randint(0, 12, int(flip.sum()))
        queries.append(q)
    queries.append(rng.randint(100, 200, (12, 100)).astype(np.uint8))     # no match
    return entries, queries


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mnemophonix_b200 import capi, dist as mdist
    from oracle.oracle import Oracle
    o = Oracle()
    entries, queries = _make_corpus()
    counts = [len(s) for s in entries]
    cuts = mdist.entry_ranges(counts, world)
    a, b = cuts[rank], cuts[rank + 1]
    shard = o.make_db(entries[a:b])
    total = sum(counts)
    qstart = np.zeros(len(queries) + 1, np.uint32)
    np.cumsum([len(q) for q in queries], out=qstart[1:])
    cands, lasts = [], []
    for qi, q in enumerate(queries):
        sc, last = o.search_shard(shard, total // 2, q)
        for e in np.nonzero(sc[:, 1])[0]:
            cands.append((qi, a + int(e), float(sc[e, 0]), int(sc[e, 1])))
        last = last.copy()
        last[:, 1] += a                                                   # global entry index
        lasts.append(last)
    cands = np.array(cands, capi.CAND_DTYPE) if cands else np.zeros(0, capi.CAND_DTYPE)
    last = np.zeros(int(qstart[-1]), capi.LAST_DTYPE)
    flat = np.concatenate(lasts)
    for k, name in enumerate(("has", "entry", "sig", "hits", "score")):
        last[name] = flat[:, k]
    res = mdist.merge_shards(cands, last, qstart, len(entries), device="cpu")
    # index-side bookkeeping: every rank computes the same assignment, shares nothing else
    assign = mdist.shard_tracks([30, 5, 240, 240, 7200, 12, 12, 600], world)
    if rank == 0:
        out.put((res, assign, cuts))
    dist.barrier()
    dist.destroy_process_group()


class _OracleShard:
    """Stand-in for capi.Db in the ranked protocol: probe + scoring by the oracle on this rank's shard, the per-node
    lists by the Python replay of the ranking tree (tests/ranking_ref.py)."""

    def __init__(self, oracle, sigs, table_size, entry_base):
        self.o, self.sigs, self.table_size, self.base = oracle, sigs, table_size, entry_base
        self.db = oracle.make_db(sigs) if len(sigs) else None

    def dense_begin(self, prepared):
        qs, qstart = prepared
        self.qstart = qstart
        nq = len(qstart) - 1
        n = len(self.sigs)
        self.acc = np.zeros((nq, max(n, 1), 2), np.float64)
        self.last = np.zeros((int(qstart[-1]), 5), np.uint32)
        for q in range(nq):
            if self.db is None:
                continue
            sc, last = self.o.search_shard(self.db, self.table_size, qs[qstart[q]:qstart[q + 1]])
            self.acc[q, :n] = sc
            self.last[qstart[q]:qstart[q + 1]] = last
        return self.last[:, 0].astype(np.uint8), (0, 0)

    def dense_lists(self, has_higher, node_bounds):
        from ranking_ref import node_list
        from mnemophonix_b200 import capi
        nq = len(self.qstart) - 1
        acc = self.acc.copy()
        for q in range(nq):
            for s in range(int(self.qstart[q]), int(self.qstart[q + 1])):
                has, e, _, hits, score = (int(x) for x in self.last[s])
                if has and has_higher[s] and hits >= 2 and score >= 30:      # this shard's own last group, scored after all
                    acc[q, e, 0] += score
                    acc[q, e, 1] += 1
        n_nodes = len(node_bounds) - 1
        lists = np.zeros((n_nodes, nq, 10), capi.SEL_DTYPE)
        counts = np.zeros((n_nodes, nq), np.uint32)
        for i in range(n_nodes):
            lo, n = int(node_bounds[i]) - self.base, int(node_bounds[i + 1] - node_bounds[i])
            for q in range(nq):
                s = acc[q, :, 0].astype(np.float32)
                c = acc[q, :, 1].astype(np.int32)
                lst = node_list(s, c, lo, n)
                counts[i, q] = len(lst)
                for k, (avg, nn, e, sc) in enumerate(lst):
                    lists[i, q, k] = (self.base + e, sc, nn, avg)
        return lists, counts

    def close(self):
        pass


def _ranked_worker(rank, world, port, out, tile=None):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mnemophonix_b200 import capi, dist as mdist
    from oracle.oracle import Oracle
    o = Oracle()
    entries, queries = _make_corpus()
    total = sum(len(s) for s in entries)
    cuts, _ = capi.tree_cuts(len(entries), world)
    mine = entries[int(cuts[rank]):int(cuts[rank + 1])]
    s = mdist.RankedShardedSearcher(None, mine, total // 2, len(entries), rank, world, device="cpu",
                                    db_factory=lambda sigs, ts, base: _OracleShard(o, sigs, ts, base), tile_queries=tile)
    res = s.search(capi.Db.prepare(queries))
    assert s.message_bytes > 0
    s.close()
    if rank == 0:
        out.put(res)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,tile", [(2, None), (3, None), (2, 2)])
def test_ranked_shards_over_gloo_equal_single_database(world, tile):
    """dist.RankedShardedSearcher over gloo (two all-gathers per query tile, last-group fix-up, merge of the node lists),
    the GPU work of each rank stood in for by the oracle: 2 ranks (one node each), 3 ranks (rank 2 owns two nodes), and
    2 ranks with the five queries cut into tiles of two."""
    from oracle.oracle import Oracle
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ranked_worker, args=(r, world, port, out, tile)) for r in range(world)]
    for p in procs:
        p.start()
    res = out.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    o = Oracle()
    entries, queries = _make_corpus()
    db = o.make_db(entries)
    want = []
    for q in queries:
        best, sc, _, _ = o.search(db, q)
        want.append((best, float(sc[best, 0]) if best >= 0 else 0.0, int(sc[best, 1]) if best >= 0 else 0))
    assert res == want


def _grid_worker(rank, world, port, out, db_shards):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mnemophonix_b200 import capi, dist as mdist
    from oracle.oracle import Oracle
    o = Oracle()
    entries, queries = _make_corpus()
    total = sum(len(s) for s in entries)
    shard, _group = mdist.grid_layout(world, db_shards)[rank]
    cuts, _ = capi.tree_cuts(len(entries), db_shards)
    mine = entries[int(cuts[shard]):int(cuts[shard + 1])]
    s = mdist.GridShardedSearcher(None, mine, total // 2, len(entries), rank, world, db_shards, device="cpu",
                                  db_factory=lambda sigs, ts, base: _OracleShard(o, sigs, ts, base))
    res = s.search(capi.Db.prepare(queries))
    s.close()
    out.put((rank, res.tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,db_shards", [(4, 2), (2, 1), (3, 3)])
def test_grid_of_shards_and_query_groups_over_gloo(world, db_shards):
    """dist.GridShardedSearcher: world = database shards x query groups.  4 ranks as two replicas of a two-way split
    (each pair runs the ranked-shard protocol in its own subgroup on its slice of the queries), 2 ranks as two whole
    replicas (no database collective at all), 3 ranks as one three-way split (the plain protocol); every rank ends up
    with every answer, equal to the oracle's search of the single database."""
    from oracle.oracle import Oracle
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grid_worker, args=(r, world, port, out, db_shards)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(out.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    o = Oracle()
    entries, queries = _make_corpus()
    db = o.make_db(entries)
    want = []
    for q in queries:
        best, sc, _, _ = o.search(db, q)
        want.append((best, float(sc[best, 0]) if best >= 0 else 0.0, int(sc[best, 1]) if best >= 0 else 0))
    for r in range(world):
        assert got[r] == want, r


def test_query_slices_balance_query_signatures():
    from mnemophonix_b200 import dist as mdist
    rng = np.random.RandomState(3)
    for n_groups in (1, 2, 4, 8):
        lens = rng.randint(0, 60, 37)
        qstart = np.concatenate([[0], np.cumsum(lens)])
        cuts = mdist.query_slices(qstart, n_groups)
        assert len(cuts) == n_groups + 1 and cuts[0] == 0 and cuts[-1] == 37
        assert all(a <= b for a, b in zip(cuts, cuts[1:]))
        per = [int(qstart[b] - qstart[a]) for a, b in zip(cuts, cuts[1:])]
        assert max(per) <= qstart[-1] / n_groups + lens.max()
    assert mdist.query_slices(np.zeros(1, np.uint32), 4) == [0, 0, 0, 0, 0]
    assert mdist.grid_layout(8, 2)[5] == (1, 2)


def test_two_rank_search_merge_equals_single_database():
    from oracle.oracle import Oracle
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res, assign, cuts = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    o = Oracle()
    entries, queries = _make_corpus()
    db = o.make_db(entries)
    want = []
    for q in queries:
        best, sc, _, _ = o.search(db, q)
        want.append((best, float(sc[best, 0]) if best >= 0 else 0.0, int(sc[best, 1]) if best >= 0 else 0))
    assert res == want
    assert [r[0] for r in res][:4] == [1, 4, 8, 8] and res[4][0] == -7
    assert sorted(i for lst in assign for i in lst) == list(range(8))
    assert 4 in assign[0] and len(assign[1]) == 7                        # the 2 h track alone on one rank
    assert cuts[0] == 0 and cuts[-1] == 9 and 0 < cuts[1] < 9


def test_entry_ranges_and_track_sharding_properties():
    from mnemophonix_b200 import dist as mdist
    rng = np.random.RandomState(0)
    for world in (1, 2, 4, 8):
        counts = rng.randint(0, 3000, 57).tolist()
        cuts = mdist.entry_ranges(counts, world)
        assert len(cuts) == world + 1 and cuts[0] == 0 and cuts[-1] == 57
        assert all(x <= y for x, y in zip(cuts, cuts[1:]))
        per = [sum(counts[a:b]) for a, b in zip(cuts, cuts[1:])]
        assert max(per) <= sum(counts) / world + max(counts)
        durs = rng.randint(1, 500, 40).tolist()
        assign = mdist.shard_tracks(durs, world)
        loads = [sum(durs[i] for i in lst) for lst in assign]
        assert sorted(i for lst in assign for i in lst) == list(range(40))
        assert max(loads) - min(loads) <= max(durs)
