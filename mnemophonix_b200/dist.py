"""torch.distributed plumbing for the multi-GPU paths (one process per GPU).

Index: tracks are independent units -> sharded across ranks with NO data-path collective
(SURVEY.md §8e); per-track results are gathered on the host only to restore input order.
Search: the database is split into contiguous entry ranges, one per rank; every rank probes its
shard with the GLOBAL table size (lsh.c:61) and the per-query candidate records are exchanged with
ONE all-gather (NCCL on GPU tensors; gloo on CPU tensors in the tests), after which the exact final
ranking (search.c:170-193) runs on the host.  The records are a few hundred bytes per query: the
exchange is latency-bound, which is why it is a single fixed-capacity collective.
"""
from __future__ import annotations

import numpy as np

from . import capi


# ------------------------------------------------------------------ index sharding ----

def shard_tracks(durations, world: int):
    """Longest-first greedy assignment of tracks to ranks.  Returns a list of index lists."""
    order = sorted(range(len(durations)), key=lambda i: -durations[i])
    load = [0.0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += durations[i]
    for lst in out:
        lst.sort()
    return out


def entry_ranges(sig_counts, world: int):
    """Contiguous entry ranges with roughly equal signature counts.  Returns world+1 cut points."""
    total = int(sum(sig_counts))
    cuts, acc, e = [0], 0, 0
    n = len(sig_counts)
    for r in range(1, world):
        target = total * r / world
        while e < n and acc + sig_counts[e] / 2 <= target:
            acc += sig_counts[e]
            e += 1
        cuts.append(e)
    cuts.append(n)
    for i in range(1, len(cuts)):
        cuts[i] = max(cuts[i], cuts[i - 1])
    return cuts


def index_sharded(ctx, pcms, rank: int, world: int, group=None, gather_to: int | None = 0):
    """Every rank indexes its share of `pcms` (all ranks pass the same list, or None for tracks they
    do not own).  With gather_to set, that rank receives the per-track results in input order."""
    import torch.distributed as dist
    durations = [0 if p is None else len(p) for p in pcms]
    if world > 1:
        # ranks may hold only their own tracks: agree on the durations first (host-side metadata)
        alld = [None] * world
        dist.all_gather_object(alld, durations, group=group)
        durations = [max(d[i] for d in alld) for i in range(len(pcms))]
    mine = shard_tracks(durations, world)[rank]
    sigs, status = ctx.index_pcm_batch([pcms[i] for i in mine]) if mine else ([], [])
    local = {i: (s, st) for i, s, st in zip(mine, sigs, status)}
    if world == 1 or gather_to is None:
        return local
    gathered = [None] * world if rank == gather_to else None
    dist.gather_object(local, gathered, dst=gather_to, group=group)
    if rank != gather_to:
        return None
    merged = {}
    for part in gathered:
        merged.update(part)
    return [merged[i] for i in range(len(pcms))]


# ------------------------------------------------------------------ sharded search ----

def _all_gather_bytes(buf: np.ndarray, device, group):
    """All-gather of variable-length byte buffers: counts first, then one padded all-gather."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n = torch.tensor([buf.nbytes], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    send = torch.zeros(cap, dtype=torch.uint8, device=device)
    if buf.nbytes:
        send[:buf.nbytes] = torch.from_numpy(np.frombuffer(buf.tobytes(), np.uint8).copy()).to(device)
    recv = [torch.empty(cap, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(recv, send, group=group)
    return [r[:c].cpu().numpy().tobytes() for r, c in zip(recv, counts)]


def merge_shards(cands_local: np.ndarray, last_local: np.ndarray, qstart: np.ndarray, n_entries_total: int,
                 device="cpu", group=None):
    """The exchange step: all-gather every rank's candidate and last-group records, then the exact
    final ranking on the host.  Every rank returns the same result list."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    parts_c = _all_gather_bytes(np.ascontiguousarray(cands_local, capi.CAND_DTYPE), device, group)
    parts_l = _all_gather_bytes(np.ascontiguousarray(last_local, capi.LAST_DTYPE), device, group)
    cands = np.concatenate([np.frombuffer(b, capi.CAND_DTYPE) for b in parts_c]) if parts_c else cands_local
    last = np.stack([np.frombuffer(b, capi.LAST_DTYPE) for b in parts_l])
    return capi.search_finish(cands, last, world, qstart, n_entries_total)


def search_sharded(ctx, entry_sigs, queries, rank: int, world: int, device, group=None, db=None):
    """entry_sigs: the WHOLE database as a list of [n,100] arrays (every rank passes the same list; only
    its own range is uploaded).  Returns (results, db) so the shard can be reused for more batches."""
    counts = [len(s) for s in entry_sigs]
    cuts = entry_ranges(counts, world)
    a, b = cuts[rank], cuts[rank + 1]
    if db is None:
        db = capi.Db(ctx, entry_sigs[a:b], table_size=int(sum(counts)) // 2, entry_base=a)
    cands, last, qstart, _ = db.search_shard(queries)   # queries: list, or capi.Db.prepare(list)
    if world == 1:
        return capi.search_finish(cands, None, 1, qstart, len(entry_sigs)), db
    return merge_shards(cands, last, qstart, len(entry_sigs), device=device, group=group), db


def search_sharded_push(ctx, entry_sigs, queries, rank: int, world: int, group=None, db=None, merge_rank: int = 0,
                        cap: int | None = None):
    """Sharded search with the merge fused into the kernels: every rank's search epilogue writes its candidate
    and last-group records straight into the merging rank's HBM over NVLink peer memory (CUDA IPC mapping);
    the only inter-process traffic besides that is the 64-byte handle and two barriers.  Returns
    (results on merge_rank / None elsewhere, db)."""
    import torch.distributed as dist
    counts = [len(s) for s in entry_sigs]
    cuts = entry_ranges(counts, world)
    a, b = cuts[rank], cuts[rank + 1]
    if db is None:
        db = capi.Db(ctx, entry_sigs[a:b], table_size=int(sum(counts)) // 2, entry_base=a)
    prepared = queries if isinstance(queries, tuple) else capi.Db.prepare(queries)
    qstart = prepared[1]
    nq, nqs = len(qstart) - 1, int(qstart[-1])
    cap = cap or max(4096, 64 * nq)
    if world == 1:
        res, _ = capi.Result.create(ctx, cap, 1, nqs)
        res.push(db, prepared, 0)
        out = res.finish(qstart, len(entry_sigs))
        res.close()
        return out, db
    box = [None]
    if rank == merge_rank:
        res, handle = capi.Result.create(ctx, cap, world, nqs)
        box[0] = handle
    dist.broadcast_object_list(box, src=merge_rank, group=group)
    if rank != merge_rank:
        res = capi.Result.open(ctx, box[0], cap, world, nqs)
    dist.barrier(group=group)              # the buffer is zeroed and mapped everywhere
    res.push(db, prepared, rank)           # returns once this rank's records are in the merging rank's memory
    dist.barrier(group=group)
    out = res.finish(qstart, len(entry_sigs)) if rank == merge_rank else None
    dist.barrier(group=group)              # nobody unmaps before the owner has read
    res.close()
    return out, db


class ShardedSearcher:
    """A database shard per rank plus ONE result buffer kept mapped on every rank (peer memory over NVLink),
    reused for every query batch of the same size: what a resident search service does.  Per batch the ranks
    meet at two barriers (buffer cleared / records written); there is no other inter-process traffic.

        s = ShardedSearcher(ctx, my_entry_sigs, table_size, entry_base, n_entries_total, rank, world)
        results = s.search(prepared_queries)        # list on merge_rank, None elsewhere
    """

    def __init__(self, ctx, shard_sigs, table_size: int, entry_base: int, n_entries_total: int, rank: int,
                 world: int, group=None, merge_rank: int = 0):
        self.ctx, self.rank, self.world, self.group, self.merge_rank = ctx, rank, world, group, merge_rank
        self.n_entries_total = n_entries_total
        self.db = capi.Db(ctx, shard_sigs, table_size=table_size, entry_base=entry_base)
        self.res, self.res_key = None, None

    def _barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier(group=self.group)

    def _result(self, cap: int, nqs: int):
        if self.res_key == (cap, nqs):
            return self.res
        self._drop_result()
        if self.world == 1:
            self.res, _ = capi.Result.create(self.ctx, cap, 1, nqs)
        else:
            import torch.distributed as dist
            box = [None]
            if self.rank == self.merge_rank:
                self.res, box[0] = capi.Result.create(self.ctx, cap, self.world, nqs)
            dist.broadcast_object_list(box, src=self.merge_rank, group=self.group)
            if self.rank != self.merge_rank:
                self.res = capi.Result.open(self.ctx, box[0], cap, self.world, nqs)
        self.res_key = (cap, nqs)
        return self.res

    def search(self, queries, cap: int | None = None):
        prepared = queries if isinstance(queries, tuple) else capi.Db.prepare(queries)
        qstart = prepared[1]
        nq, nqs = len(qstart) - 1, int(qstart[-1])
        res = self._result(cap or max(4096, 64 * nq), nqs)
        if self.rank == self.merge_rank:
            res.reset()
        self._barrier()                       # the buffer is cleared (and mapped everywhere)
        res.push(self.db, prepared, self.rank)
        self._barrier()                       # every shard's records are in the merging rank's memory
        return res.finish(qstart, self.n_entries_total) if self.rank == self.merge_rank else None

    def close(self):
        self._drop_result()
        self.db.close()

    def _drop_result(self):
        if self.res is None:
            return
        self._barrier()                       # nobody unmaps before the owner has read
        if self.rank != self.merge_rank:
            self.res.close()
        self._barrier()                       # the owner frees once every peer mapping is gone
        if self.rank == self.merge_rank:
            self.res.close()
        self.res, self.res_key = None, None


class RankedShardedSearcher:
    """Sharded search with the ranking done where the entries live (k_select.cu): shards sit on node boundaries of
    the reference's merge-sort tree, each rank ranks its own nodes on its GPU and contributes ten records per query
    and node.  Two small all-gathers per batch (NCCL on GPU tensors; gloo on CPU tensors in tests): one byte per
    query signature ("this shard holds a candidate"), then the lists.  Every rank returns the results.

        cuts, node_lo = capi.tree_cuts(n_entries_total, world)       # rank r owns entries cuts[r] .. cuts[r+1]-1
        s = RankedShardedSearcher(ctx, sigs_of_my_entries, table_size, n_entries_total, rank, world, device)
        results = s.search(prepared_queries)
    """

    def __init__(self, ctx, shard_sigs, table_size: int, n_entries_total: int, rank: int, world: int, device="cpu",
                 group=None, db_factory=None):
        """db_factory(shard_sigs, table_size, entry_base) -> object with dense_begin / dense_lists / close; default:
        capi.Db on ctx (the CPU tests pass a stand-in so that the plumbing runs without a GPU)."""
        self.ctx, self.rank, self.world, self.device, self.group = ctx, rank, world, device, group
        self.cuts, self.node_lo = capi.tree_cuts(n_entries_total, world)
        n_nodes = len(self.node_lo) - 1
        self.node_range = (rank * n_nodes // world, (rank + 1) * n_nodes // world)     # my nodes
        a, b = int(self.cuts[rank]), int(self.cuts[rank + 1])
        assert len(shard_sigs) == b - a, "shard does not match capi.tree_cuts"
        self.db = db_factory(shard_sigs, table_size, a) if db_factory else \
            capi.Db(ctx, shard_sigs, table_size=table_size, entry_base=a)

    def _all_gather(self, arr: np.ndarray):
        """Equal-shaped numpy arrays from every rank, in rank order."""
        if self.world == 1:
            return [arr]
        import torch
        import torch.distributed as dist
        t = torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1)).to(self.device)
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t, group=self.group)
        return [o.cpu().numpy().view(arr.dtype).reshape(arr.shape) for o in out]

    def search(self, queries):
        prepared = queries if isinstance(queries, tuple) else capi.Db.prepare(queries)
        has, _ = self.db.dense_begin(prepared)
        all_has = self._all_gather(has)                                   # the exchange step: 1 byte per query signature
        higher = np.zeros_like(has)
        for r in range(self.rank + 1, self.world):
            higher |= all_has[r]
        n0, n1 = self.node_range
        lists, counts = self.db.dense_lists(higher, self.node_lo[n0:n1 + 1])
        # ranks may own different numbers of nodes (world not a power of two): pad to the largest for the all-gather
        n_nodes = len(self.node_lo) - 1
        per = max((r + 1) * n_nodes // self.world - r * n_nodes // self.world for r in range(self.world))
        nq = counts.shape[1]
        pl = np.zeros((per, nq, 10), capi.SEL_DTYPE)
        pc = np.zeros((per, nq), np.uint32)
        pl[:n1 - n0], pc[:n1 - n0] = lists, counts
        gl, gc = self._all_gather(pl), self._all_gather(pc)
        all_l = np.concatenate([gl[r][:(r + 1) * n_nodes // self.world - r * n_nodes // self.world] for r in range(self.world)])
        all_c = np.concatenate([gc[r][:(r + 1) * n_nodes // self.world - r * n_nodes // self.world] for r in range(self.world)])
        return capi.select_merge(all_l, all_c)

    def close(self):
        self.db.close()
