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
