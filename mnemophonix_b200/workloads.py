"""Benchmark workloads shared by bench.py and tools/ (measurement plumbing, not product code).

c4_search: BASELINE.json configs[3] — search N five-second excerpts against an N-track database sharded across the
ranks of one box (one process per GPU), with the parity checks SURVEY.md §8d asks for:
  * every answer of the sharded search equals the single-GPU answer (the whole database gathered onto rank 0);
  * a sample of >= 64 queries is handed to the caller's checker together with the dense per-entry (score, n_matches)
    rows of those queries — the single-GPU rows and the sharded rows after the cross-shard last-group fix-up — so that
    bench.py can compare answers with the unmodified reference and rows with the oracle's search.c:55-59 rows.
"""
from __future__ import annotations

import os
import time

import numpy as np

from . import capi, dist as mdist, synth

RATE = 44100


def _pad_gather(dist, lists, per, dev, torch):
    """All-gather of per-rank lists of [n_i, 100] u8 arrays (ragged): returns (counts [world*per], data [world*per, max, 100])."""
    world = dist.get_world_size() if dist is not None else 1
    cnt = torch.zeros(per, dtype=torch.int64, device=dev)
    cnt[:len(lists)] = torch.tensor([len(x) for x in lists], dtype=torch.int64)
    mx = torch.tensor([max([len(x) for x in lists] + [1])], dtype=torch.int64, device=dev)
    if dist is not None:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    mx = int(mx.item())
    mine = torch.zeros((per, mx, 100), dtype=torch.uint8)
    for i, x in enumerate(lists):
        mine[i, :len(x)] = torch.from_numpy(np.ascontiguousarray(x))
    mine = mine.to(dev)
    if dist is None:
        return cnt.cpu().numpy(), mine.cpu().numpy()
    allc = torch.empty(world * per, dtype=torch.int64, device=dev)
    alld = torch.empty((world * per, mx, 100), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(allc, cnt)
    dist.all_gather_into_tensor(alld, mine)
    return allc.cpu().numpy(), alld.cpu().numpy()


def c4_search(ctx, stream, dev, rank: int, world: int, dist, n_tracks: int = 10000, seconds: float = 30.0, chunk: int = 50,
              steps: int = 3, ref_queries: int = 64, n_queries: int = 0, host_threads: int = 0, checker=None):
    """Returns the result dict on rank 0 (None elsewhere).  dist: torch.distributed (initialised, NCCL) or None.
    checker(entries, queries, pick, got, single_score_rows, single_n_rows, sharded_rows, threads) -> dict: the caller's
    comparison with the reference / oracle on the sampled queries (bench.py:c4_checker; this package never touches them)."""
    import torch
    cuts, _ = capi.tree_cuts(n_tracks, world)        # shards on node boundaries of the ranking tree
    a, b = int(cuts[rank]), int(cuts[rank + 1])
    per = int(max(cuts[1:] - cuts[:-1]))
    q_off = int(10.0 * RATE) + 1234 if seconds < 70 else int(60.0 * RATE) + 1234     # not hop-aligned
    q_len = 5 * RATE
    my_sigs, my_queries = [], []
    t_synth = t_index = 0.0
    for c0 in range(a, b, chunk):
        ids = list(range(c0, min(b, c0 + chunk)))
        t0 = time.perf_counter()
        pcm = synth.synth_tracks_torch(ids, seconds, dev)
        host = torch.empty(pcm.shape, dtype=torch.int16).pin_memory()
        host.copy_(pcm)
        torch.cuda.synchronize()
        t_synth += time.perf_counter() - t0
        t0 = time.perf_counter()
        arr = host.numpy()
        sigs, st = ctx.index_pcm_batch([arr[i].reshape(-1, 1) for i in range(len(ids))])
        qs, qst = ctx.index_pcm_batch([np.ascontiguousarray(arr[i, q_off:q_off + q_len]).reshape(-1, 1) for i in range(len(ids))])
        t_index += time.perf_counter() - t0
        assert not any(st) and not any(qst)
        my_sigs += sigs
        my_queries += qs
        del pcm, host
    # every rank needs every query (queries are replicated to all shards)
    qcnt, qdat = _pad_gather(dist, my_queries, per, dev, torch)
    own = [int(cuts[r + 1] - cuts[r]) for r in range(world)]
    order = [r * per + i for r in range(world) for i in range(own[r])]
    n_q = n_queries or n_tracks
    queries = [qdat[s, :qcnt[s]] for s in order][:n_q]
    prepared = capi.Db.prepare(queries, ctx)          # page-locked: the upload runs at the H2D rate
    n_qsig = int(prepared[1][-1])
    n_local = torch.tensor([sum(len(s) for s in my_sigs)], dtype=torch.int64, device=dev)
    if dist is not None:
        dist.all_reduce(n_local)
    total_sigs = int(n_local.item())

    t0 = time.perf_counter()
    searcher = mdist.RankedShardedSearcher(ctx, my_sigs, total_sigs // 2, n_tracks, rank, world, device=dev, stream=stream)
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t0

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    res = searcher.search(prepared)                      # warm-up
    times = []
    for _ in range(steps):
        barrier()
        t0 = time.perf_counter()
        res = searcher.search(prepared, as_array=True)   # 12 bytes per query in host memory
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if dist is not None:
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        times.append(dt)
    dt = float(np.mean(times))
    res = res.tolist()
    searcher.profile = True                              # one more batch with CUDA events between the phases
    barrier()
    searcher.search(prepared, as_array=True)
    phases = dict(searcher.phase_ms)
    searcher.profile = False
    msg_bytes = searcher.message_bytes
    L, D = searcher.db.dense_stats() if n_q <= searcher.tile else (0, 0)
    LD = torch.tensor([L, D], dtype=torch.int64, device=dev)
    if dist is not None:
        dist.all_reduce(LD)
    L, D = int(LD[0]), int(LD[1])

    # ---- parity ----
    pick = np.unique(np.linspace(0, n_q - 1, min(ref_queries, n_q)).astype(int)) if ref_queries else np.zeros(0, int)
    sample = capi.Db.prepare([queries[i] for i in pick]) if len(pick) else None
    sharded_rows = None
    if sample is not None:
        searcher.search(sample)                          # collective: leaves the sample's accumulators on every shard
        ne = per
        rows = torch.zeros((len(pick), ne, 2), dtype=torch.int32)
        for k in range(len(pick)):
            sc, nm = searcher.db.dense_read(k)
            rows[k, :len(sc), 0] = torch.from_numpy(sc.astype(np.int32))
            rows[k, :len(nm), 1] = torch.from_numpy(nm.astype(np.int32))
        if dist is not None:
            allr = torch.empty((world,) + tuple(rows.shape), dtype=torch.int32, device=dev)
            dist.all_gather_into_tensor(allr, rows.to(dev))
            allr = allr.cpu().numpy()
            sharded_rows = np.concatenate([allr[r][:, :own[r]] for r in range(world)], axis=1)
        else:
            sharded_rows = rows.numpy()[:, :own[0]]
    scnt, sdat = _pad_gather(dist, my_sigs, per, dev, torch)    # the whole database, for the checks on rank 0
    out = None
    if rank == 0:
        entries = [sdat[s, :scnt[s]] for s in order]
        got = [r[0] for r in res]
        out = {"workload": f"{n_q} five-second excerpts x {n_tracks}-track database ({seconds:g} s tracks from GPU-synthesised PCM, "
                           f"{total_sigs} signatures) sharded over {world} GPU(s) on ranking-tree nodes; host query signatures in, "
                           "answers out; two all-gathers on device tensors per query tile, merge on the GPU",
               "n_gpus": world, "queries_per_s": n_q / dt, "s_per_batch": dt, "db_signatures": total_sigs,
               "query_signatures": n_qsig, "lsh_build_s_per_shard": t_build, "index_s_this_rank": t_index,
               "synth_s_this_rank": t_synth, "query_tiles": -(-n_q // searcher.tile), "phase_ms": phases,
               "nccl_message_bytes_per_batch_per_rank": int(msg_bytes),
               "excerpts_identified": int(sum(int(g == i) for i, g in enumerate(got))),
               "probed_nodes_per_query_signature": L / max(n_qsig, 1), "deep_checks_per_query_signature": D / max(n_qsig, 1),
               "algorithmic_GBps": (n_qsig * 300 + 4 * L + 100 * D + 16 * n_q) / dt / 1e9}
        # (a) the same queries against the whole database on ONE GPU
        t0 = time.perf_counter()
        single = capi.Db(ctx, entries)
        out["single_gpu_lsh_build_s"] = time.perf_counter() - t0
        single.search(prepared)
        t0 = time.perf_counter()
        want, _ = single.search(prepared)
        out["single_gpu_queries_per_s"] = n_q / (time.perf_counter() - t0)
        out["answers_equal_single_gpu"] = int(sum(int(tuple(x) == tuple(y)) for x, y in zip(res, want)))
        out["identical_to_single_gpu"] = out["answers_equal_single_gpu"] == n_q
        if sample is not None and checker is not None:
            s_sc, s_nm = single.scores(sample)
            try:
                out.update(checker(entries, queries, pick, got, s_sc, s_nm, sharded_rows, host_threads or min(32, os.cpu_count() or 1)))
            except Exception as exc:   # the checker is optional here
                out["reference_check_error"] = repr(exc)
        single.close()
    searcher.close()
    return out
