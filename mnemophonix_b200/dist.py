"""torch.distributed plumbing for the multi-GPU paths (one process per GPU).

Index: tracks are independent units -> sharded across ranks with NO data-path collective
(SURVEY.md §8e); per-track results are gathered on the host only to restore input order.
Search: RankedShardedSearcher — the database is split on node boundaries of the reference's ranking
tree, one contiguous entry range per rank; every rank probes its shard with the GLOBAL table size
(lsh.c:61), ranks its own nodes on its GPU, and two small all-gathers per query tile (NCCL on device
tensors, no host bounce) carry one byte per query signature and ten records per query and node; the
top levels of the tree are merged on the GPU.  GridShardedSearcher lays the ranks out as database shards x query
groups (shard for capacity, replicate for throughput): each group of S ranks is one RankedShardedSearcher over an
S-way split database and answers its own slice of the query batch; one more all-gather of 12 bytes per query
returns every answer to every rank.  search_sharded / merge_shards are the older record form (arbitrary cuts, host
merge), kept as a diagnostic.
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


class RankedShardedSearcher:
    """Sharded search with the ranking done where the entries live (k_select.cu): shards sit on node boundaries of
    the reference's merge-sort tree, each rank ranks its own nodes on its GPU and contributes ten records per query
    and node.  Two small all-gathers per query tile, on device tensors and on the context's stream (NCCL; gloo on
    CPU tensors in the tests, where a stand-in replaces the GPU probe): one byte per query signature ("this shard
    holds a candidate"), then the lists; the top levels of the tree are merged on every rank's GPU and 12 bytes per
    query come back to the host.  Every rank returns the results.

        cuts, node_lo = capi.tree_cuts(n_entries_total, world)       # rank r owns entries cuts[r] .. cuts[r+1]-1
        s = RankedShardedSearcher(ctx, sigs_of_my_entries, table_size, n_entries_total, rank, world, device,
                                  stream=torch_stream_the_context_runs_on)
        results = s.search(prepared_queries)
    """

    PHASES = ("probe", "has_gather", "fixup_rank", "list_gather", "merge")

    def __init__(self, ctx, shard_sigs, table_size: int, n_entries_total: int, rank: int, world: int, device="cpu",
                 group=None, db_factory=None, stream=None, tile_queries: int | None = None):
        """db_factory(shard_sigs, table_size, entry_base) -> object with dense_begin / dense_lists / close (host
        arrays): the CPU tests pass an oracle-backed stand-in so that the collectives run without a GPU.  stream: the
        torch.cuda.Stream whose handle the context was created with (None: the device is synchronised around the
        collectives instead)."""
        import os
        self.ctx, self.rank, self.world, self.device, self.group, self.stream = ctx, rank, world, device, group, stream
        self.cuts, self.node_lo = capi.tree_cuts(n_entries_total, world)
        self.n_nodes = len(self.node_lo) - 1
        own = [(r + 1) * self.n_nodes // world - r * self.n_nodes // world for r in range(world)]
        self.per = max(own)                                                                 # nodes per shard block
        self.node_range = (rank * self.n_nodes // world, (rank + 1) * self.n_nodes // world)     # my nodes
        a, b = int(self.cuts[rank]), int(self.cuts[rank + 1])
        assert len(shard_sigs) == b - a, "shard does not match capi.tree_cuts"
        self.on_gpu = db_factory is None
        self.db = db_factory(shard_sigs, table_size, a) if db_factory else \
            capi.Db(ctx, shard_sigs, table_size=table_size, entry_base=a)
        # every rank must cut the batch into the same tiles: sized for the largest shard
        budget = int(os.environ.get("MNEMO_ACC_BYTES", 1 << 30))
        widest = max(int(self.cuts[r + 1] - self.cuts[r]) for r in range(world))
        self.tile = tile_queries or max(1, budget // (8 * max(widest, 1)))
        self.phase_ms = dict.fromkeys(self.PHASES, 0.0)
        self.profile = False
        self.message_bytes = 0          # bytes this rank contributed to the collectives of the last search()

    # -- collectives on flat uint8 tensors --
    def _gather(self, t):
        import torch
        import torch.distributed as dist
        out = torch.empty(self.world * t.numel(), dtype=torch.uint8, device=t.device)
        dist.all_gather_into_tensor(out, t, group=self.group)
        self.message_bytes += t.numel()
        return out

    def search(self, queries, as_array: bool = False):
        """[(entry, score, n_matches)] per query; as_array (GPU path): one capi.MATCH_DTYPE array instead — a list of
        10 000 tuples takes Python longer to build than the merge kernel runs."""
        qs, qstart = queries if isinstance(queries, tuple) else capi.Db.prepare(queries)
        nq = len(qstart) - 1
        self.message_bytes = 0
        if self.profile:
            self.phase_ms = dict.fromkeys(self.PHASES, 0.0)
        out = []
        for q0 in range(0, max(nq, 1), self.tile):
            q1 = min(nq, q0 + self.tile)
            s0, s1 = int(qstart[q0]), int(qstart[q1])
            sub = (qs[s0:s1], np.ascontiguousarray(qstart[q0:q1 + 1] - qstart[q0]).astype(np.uint32))
            out.append(self._tile_gpu(sub, as_array) if self.on_gpu else self._tile_host(sub))
        if as_array and self.on_gpu:
            return np.concatenate(out) if out else np.zeros(0, capi.MATCH_DTYPE)
        return [r for tile in out for r in tile]

    # -- one tile, device-resident (production) --
    def _tile_gpu(self, sub, as_array=False):
        import contextlib
        import torch
        qs, qstart = sub
        nq, nqs = len(qstart) - 1, int(qstart[-1])
        if nq == 0:
            return np.zeros(0, capi.MATCH_DTYPE) if as_array else []
        n0, n1 = self.node_range
        dev = self.device
        marks = []

        def mark():
            if self.profile:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(self.stream if self.stream is not None else torch.cuda.current_stream())
                marks.append(ev)

        def fence():       # without a shared stream the collectives are ordered by device synchronisation
            if self.stream is None:
                torch.cuda.synchronize()

        scope = torch.cuda.stream(self.stream) if self.stream is not None else contextlib.nullcontext()
        with scope:
            mark()
            has = torch.empty(max(nqs, 1), dtype=torch.uint8, device=dev)
            self.db.dense_begin_dev(sub, has.data_ptr())
            fence()
            mark()
            all_has = self._gather(has) if self.world > 1 else has
            fence()
            mark()
            block = torch.zeros(capi.shard_record_bytes(self.per, nq), dtype=torch.uint8, device=dev)
            lists_ptr = block.data_ptr()
            counts_ptr = lists_ptr + self.per * nq * 10 * capi.SEL_DTYPE.itemsize
            self.db.dense_lists_dev(all_has.data_ptr(), self.world, self.rank, self.node_lo[n0:n1 + 1], lists_ptr, counts_ptr)
            fence()
            mark()
            gathered = self._gather(block) if self.world > 1 else block
            fence()
            mark()
            res = capi.select_merge_dev(self.ctx, gathered.data_ptr(), self.world, self.n_nodes, self.per, nq, as_array)   # synchronises
            mark()
        if self.profile:
            torch.cuda.synchronize()
            for name, e0, e1 in zip(self.PHASES, marks[:-1], marks[1:]):
                self.phase_ms[name] += e0.elapsed_time(e1)
        return res

    # -- one tile through the host-pointer forms (tests: gloo on CPU tensors, stand-in probe) --
    def _tile_host(self, sub):
        import torch
        qs, qstart = sub
        nq, nqs = len(qstart) - 1, int(qstart[-1])
        if nq == 0:
            return []
        has, _ = self.db.dense_begin(sub)
        has = np.ascontiguousarray(has, np.uint8)
        if self.world > 1:
            all_has = self._gather(torch.from_numpy(has.copy())).numpy().reshape(self.world, -1)
        else:
            all_has = has[None, :]
        higher = np.zeros(nqs, np.uint8)
        for r in range(self.rank + 1, self.world):
            higher |= all_has[r][:nqs]
        n0, n1 = self.node_range
        lists, counts = self.db.dense_lists(higher, self.node_lo[n0:n1 + 1])
        pl = np.zeros((self.per, nq, 10), capi.SEL_DTYPE)
        pc = np.zeros((self.per, nq), np.uint32)
        pl[:n1 - n0], pc[:n1 - n0] = lists, counts
        block = np.concatenate([pl.view(np.uint8).reshape(-1), pc.view(np.uint8).reshape(-1)])
        if self.world > 1:
            gathered = self._gather(torch.from_numpy(block)).numpy().reshape(self.world, -1)
        else:
            gathered = block[None, :]
        lb = pl.nbytes
        all_l, all_c = [], []
        for r in range(self.world):
            own = (r + 1) * self.n_nodes // self.world - r * self.n_nodes // self.world
            all_l.append(gathered[r][:lb].view(capi.SEL_DTYPE).reshape(self.per, nq, 10)[:own])
            all_c.append(gathered[r][lb:].view(np.uint32).reshape(self.per, nq)[:own])
        return capi.select_merge(np.concatenate(all_l), np.concatenate(all_c))

    def close(self):
        self.db.close()


def grid_layout(world: int, db_shards: int):
    """rank -> (shard index, query group) for a world laid out as db_shards x query groups: the ranks of one group are
    consecutive (rank = group * db_shards + shard), so a group sits on neighbouring GPUs."""
    assert db_shards >= 1 and world % db_shards == 0, "the world must be a whole number of database replicas"
    return [(r % db_shards, r // db_shards) for r in range(world)]


def query_slices(qstart, n_groups: int):
    """Contiguous slices of the query batch with roughly equal numbers of query SIGNATURES (the unit of probe work):
    n_groups + 1 cut points into the queries."""
    qstart = np.asarray(qstart, np.int64)
    nq, total = len(qstart) - 1, int(qstart[-1])
    cuts = [0]
    for g in range(1, n_groups):
        cuts.append(max(cuts[-1], min(nq, int(np.searchsorted(qstart, total * g / n_groups, side="left")))))
    cuts.append(nq)
    return cuts


class GridShardedSearcher:
    """world = db_shards x query groups.  The database is split db_shards ways (as few as its size needs: every
    shard costs each query signature a bucket lookup, its list walks and a share of two all-gathers, so a batch
    strong-scales poorly over shards) and each replica of that split — one group of db_shards ranks
    running the RankedShardedSearcher protocol among themselves over NCCL — answers 1 / groups of the query batch.  A
    last all-gather (12 bytes per query) hands every rank every answer.

        shard, group = dist.grid_layout(world, db_shards)[rank]
        cuts, _ = capi.tree_cuts(n_entries_total, db_shards)          # this rank holds entries cuts[shard] .. cuts[shard+1]-1
        g = GridShardedSearcher(ctx, sigs_of_those_entries, table_size, n_entries_total, rank, world, db_shards, device, stream=st)
        answers = g.search(prepared_queries)                          # capi.MATCH_DTYPE array, the same on every rank
    """

    def __init__(self, ctx, shard_sigs, table_size: int, n_entries_total: int, rank: int, world: int, db_shards: int,
                 device="cpu", db_factory=None, stream=None, tile_queries: int | None = None):
        import torch.distributed as dist
        self.rank, self.world, self.db_shards, self.device, self.stream = rank, world, db_shards, device, stream
        self.n_groups = world // db_shards
        self.shard, self.group_index = grid_layout(world, db_shards)[rank]
        sub = None
        if world > 1 and db_shards < world:
            # every rank creates every group, in the same order (torch.distributed.new_group is collective)
            for g in range(self.n_groups):
                grp = dist.new_group(list(range(g * db_shards, (g + 1) * db_shards))) if db_shards > 1 else None
                if g == self.group_index:
                    sub = grp
        self.inner = RankedShardedSearcher(ctx, shard_sigs, table_size, n_entries_total, self.shard, db_shards, device=device,
                                           group=sub, db_factory=db_factory, stream=stream, tile_queries=tile_queries)
        self.message_bytes = 0

    @property
    def db(self):
        return self.inner.db

    def search(self, queries):
        import torch
        import torch.distributed as dist
        qs, qstart = queries if isinstance(queries, tuple) else capi.Db.prepare(queries)
        nq = len(qstart) - 1
        cuts = query_slices(qstart, self.n_groups)
        q0, q1 = cuts[self.group_index], cuts[self.group_index + 1]
        s0, s1 = int(qstart[q0]), int(qstart[q1])
        sub = (qs[s0:s1], np.ascontiguousarray(np.asarray(qstart[q0:q1 + 1], np.int64) - int(qstart[q0])).astype(np.uint32))
        if self.inner.on_gpu:
            mine = self.inner.search(sub, as_array=True)
        else:
            mine = np.array([tuple(r) for r in self.inner.search(sub)], capi.MATCH_DTYPE).reshape(-1)
        self.message_bytes = self.inner.message_bytes
        if self.n_groups == 1:
            return mine
        per = max(b - a for a, b in zip(cuts, cuts[1:]))
        item = capi.MATCH_DTYPE.itemsize
        send = np.zeros(max(per, 1) * item, np.uint8)
        send[:len(mine) * item] = mine.view(np.uint8).reshape(-1)
        t = torch.from_numpy(send).to(self.device)
        out = torch.empty(self.world * t.numel(), dtype=torch.uint8, device=t.device)
        scope = torch.cuda.stream(self.stream) if (self.stream is not None and self.inner.on_gpu) else None
        if scope is not None:
            with scope:
                dist.all_gather_into_tensor(out, t)
        else:
            dist.all_gather_into_tensor(out, t)
        self.message_bytes += t.numel()
        if self.stream is not None and self.inner.on_gpu:
            self.stream.synchronize()
        allb = out.cpu().numpy().reshape(self.world, -1)
        parts = [allb[g * self.db_shards][: (cuts[g + 1] - cuts[g]) * item].view(capi.MATCH_DTYPE) for g in range(self.n_groups)]
        return np.concatenate(parts) if nq else np.zeros(0, capi.MATCH_DTYPE)

    def close(self):
        self.inner.close()
