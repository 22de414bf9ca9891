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
              steps: int = 5, ref_queries: int = 64, n_queries: int = 0, host_threads: int = 0, checker=None,
              grid_db_shards: int = 2):
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

    res = searcher.search(prepared)                      # warm-up (twice: scratch buffers grow on the first batch)
    searcher.search(prepared, as_array=True)
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
    dt = float(np.median(times))          # every batch is listed (s_per_batch_each)
    batch_times = [float(x) for x in times]
    res = res.tolist()
    searcher.profile = True                              # one more batch with CUDA events between the phases
    barrier()
    t0 = time.perf_counter()
    searcher.search(prepared, as_array=True)
    phases = dict(searcher.phase_ms)
    phases["host_wall_of_this_batch"] = (time.perf_counter() - t0) * 1e3
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
    searcher.close()
    entries = [sdat[s, :scnt[s]] for s in order]
    # the same batch on a grid of database shards x query groups (dist.GridShardedSearcher): the database split
    # grid_db_shards ways, world / grid_db_shards replicas of that split, each answering its slice of the queries
    grid = None
    if grid_db_shards and world > grid_db_shards and world % grid_db_shards == 0:
        shard, _ = mdist.grid_layout(world, grid_db_shards)[rank]
        gcuts, _ = capi.tree_cuts(n_tracks, grid_db_shards)
        t0 = time.perf_counter()
        gs = mdist.GridShardedSearcher(ctx, entries[int(gcuts[shard]):int(gcuts[shard + 1])], total_sigs // 2, n_tracks, rank, world,
                                       grid_db_shards, device=dev, stream=stream)
        torch.cuda.synchronize()
        g_build = time.perf_counter() - t0
        gres = gs.search(prepared)
        gs.search(prepared)
        gtimes = []
        for _ in range(steps):
            barrier()
            t0 = time.perf_counter()
            gres = gs.search(prepared)
            torch.cuda.synchronize()
            gdt = time.perf_counter() - t0
            tt = torch.tensor([gdt], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            gtimes.append(float(tt.item()))
        grid = {"layout": f"{grid_db_shards} database shards x {world // grid_db_shards} query groups",
                "queries_per_s": n_q / float(np.median(gtimes)), "s_per_batch": float(np.median(gtimes)), "s_per_batch_each": gtimes,
                "lsh_build_s_per_shard": g_build, "nccl_message_bytes_per_batch_per_rank": int(gs.message_bytes)}
        gs.close()
    out = None
    if rank == 0:
        got = [r[0] for r in res]
        out = {"workload": f"{n_q} five-second excerpts x {n_tracks}-track database ({seconds:g} s tracks from GPU-synthesised PCM, "
                           f"{total_sigs} signatures) sharded over {world} GPU(s) on ranking-tree nodes; host query signatures in, "
                           "answers out; two all-gathers on device tensors per query tile, merge on the GPU",
               "n_gpus": world, "queries_per_s": n_q / dt, "s_per_batch": dt, "s_per_batch_each": batch_times, "db_signatures": total_sigs,
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
        if grid is not None:
            grid["answers_equal_single_gpu"] = int(sum(int(tuple(x) == tuple(y)) for x, y in zip(gres.tolist(), want)))
            grid["identical_to_single_gpu"] = grid["answers_equal_single_gpu"] == n_q
            grid["speedup_vs_single_gpu"] = grid["queries_per_s"] / out["single_gpu_queries_per_s"]
            out["grid"] = grid
        if sample is not None and checker is not None:
            s_sc, s_nm = single.scores(sample)
            try:
                out.update(checker(entries, queries, pick, got, s_sc, s_nm, sharded_rows, host_threads or min(32, os.cpu_count() or 1)))
            except Exception as exc:   # the checker is optional here
                out["reference_check_error"] = repr(exc)
        single.close()
    return out


def stream_index(local_rank: int, dev, rank: int, world: int, dist, n_tracks: int, seconds: float = 240.0,
                 per_batch: int = 64, distinct: int = 13, check_every: int = 1, checker=None, label: str = "",
                 passes: int = 1):
    """BASELINE.json configs[2] (130 four-minute tracks, one pass) and configs[4] (1000 h = 15 000 tracks) — bulk
    indexing of a corpus of `n_tracks` mono tracks of `seconds` each, dealt round-robin to the ranks (independent
    units: no data-path collective).  The corpus does not fit host memory at 1000 h (317 GB of PCM), so its PCM is
    cycled from a pinned ring of `distinct` synthetic tracks; every batch's PCM still crosses PCIe
    and every batch's signatures come back to pinned host memory inside the timed region.  Each rank keeps two
    batches in flight on two streams (batch k+1's upload overlaps batch k's kernels and batch k-1's download).

    Self-check, size-independent: every occurrence of ring track r must produce the signatures (count and CRC-32) that
    r produced when indexed alone before the timed region.  checker(pcm [n, 1] int16, signatures) -> bool (bench.py:
    the CPU oracle on ring track 0) pins that single-track run.  Returns the result dict on rank 0, None elsewhere."""
    import zlib
    import torch
    mine = list(range(rank, n_tracks, world)) * passes     # a short corpus is indexed `passes` times back to back
    ids = [900000 + i for i in range(distinct)]
    ring = []
    for c0 in range(0, distinct, 4):
        pcm = synth.synth_tracks_torch(ids[c0:c0 + 4], seconds, dev)
        for k in range(pcm.shape[0]):
            h = torch.empty((pcm.shape[1], 1), dtype=torch.int16).pin_memory()
            h.copy_(pcm[k].reshape(-1, 1))
            ring.append(h)
        del pcm
    torch.cuda.synchronize()
    n_frames = ring[0].shape[0]
    lanes = []
    for _ in range(2):
        st = torch.cuda.Stream()
        lanes.append({"stream": st, "ctx": capi.Context(local_rank, st.cuda_stream), "batches": {}})
    gold_crc, gold_n, first = [], [], None
    for r in range(distinct):
        sigs, status = lanes[0]["ctx"].index_pcm_batch([ring[r].numpy()])
        assert status == [0]
        if r == 0:
            first = sigs[0].copy()
        gold_crc.append(zlib.crc32(sigs[0].tobytes()))
        gold_n.append(len(sigs[0]))

    def get_batch(lane, n):
        d = lanes[lane]["batches"]
        if n not in d:
            b = lanes[lane]["ctx"].batch([(ring[0].data_ptr(), n_frames, 1)] * n)
            d[n] = (b, torch.empty((max(b.cap, 1), 100), dtype=torch.uint8).pin_memory())
        return d[n]

    def stream(track_ids, every):
        jobs = [track_ids[i:i + per_batch] for i in range(0, len(track_ids), per_batch)]
        tot = {"sigs": 0, "checked": 0, "bad": 0}

        def collect(p):
            b, out, tids, j = p
            _, n, status = b.download(out_ptr=out.data_ptr())
            tot["sigs"] += int(n.sum())
            if any(int(s) for s in status):
                tot["bad"] += 1
            if every and j % every == 0:
                host = out.numpy()
                off = 0
                for t, cnt in zip(tids, n):
                    cnt = int(cnt)
                    r = t % distinct
                    if cnt != gold_n[r] or zlib.crc32(host[off:off + cnt].tobytes()) != gold_crc[r]:
                        tot["bad"] += 1
                    off += cnt
                    tot["checked"] += 1

        pending = None
        for j, tids in enumerate(jobs):
            b, out = get_batch(j % 2, len(tids))
            b.upload([(ring[t % distinct].data_ptr(), n_frames, 1) for t in tids])
            b.run()
            if pending is not None:
                collect(pending)
            pending = (b, out, tids, j)
        if pending is not None:
            collect(pending)
        return tot

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    stream(mine[: 3 * per_batch], 1)          # warm-up: arenas and pinned outputs allocated, both lanes used
    barrier()
    t0 = time.perf_counter()
    tot = stream(mine, check_every)
    barrier()
    dt = time.perf_counter() - t0
    # the copies alone, same ranks, same buffers: what the host side allows
    b, _ = get_batch(0, min(per_batch, max(len(mine), 1)))
    reps = max(2, min(8, len(mine) // max(per_batch, 1)))
    barrier()
    t1 = time.perf_counter()
    for _ in range(reps):
        b.upload()
    b.sync()
    copy_s = time.perf_counter() - t1
    copy_bytes = reps * b.n * n_frames * 2
    t = torch.tensor([dt, float(tot["sigs"]), float(tot["checked"]), float(tot["bad"]), float(len(mine)), float(copy_bytes), copy_s],
                     dtype=torch.float64, device=dev)
    if dist is not None:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t[0], t[6] = tmax[0], tmax[6]
    dt, sigs, checked, bad, n_done = float(t[0]), int(t[1]), int(t[2]), int(t[3]), int(t[4])
    ceil = float(t[5]) / float(t[6])          # all ranks' bytes over the slowest rank's time: the copies ran at once
    out = None
    if rank == 0:
        hours = n_done * seconds / 3600.0
        pcm_bytes = n_done * n_frames * 2
        out = {"workload": f"{label}index {n_done // passes} mono tracks of {seconds:.0f} s ({hours / passes:.1f} h, {pcm_bytes / passes / 1e9:.1f} GB of PCM{f', {passes} passes' if passes > 1 else ''}; cycled "
                           f"from a pinned ring of {distinct} synthetic tracks) dealt round-robin to {world} GPU(s), batches of "
                           f"{per_batch}, two batches in flight per GPU; host PCM in, signatures back in pinned host memory",
               "n_gpus": world, "passes": passes, "audio_hours": hours, "seconds": dt, "audio_hours_per_s": hours / dt,
               "h2d_GBps_aggregate": pcm_bytes / dt / 1e9, "h2d_ceiling_GBps_aggregate": ceil / 1e9,
               "frac_of_h2d_ceiling": (pcm_bytes / dt) / ceil if ceil > 0 else None,
               "signatures": sigs, "signature_db_text_MB_per_pass": sigs * 201 / 1e6 / passes, "d2h_bytes": sigs * 100,
               "tracks_crc_checked": checked, "tracks_differing_from_single_track_run": bad}
        if checker is not None:
            try:
                out["ring_track_equals_oracle"] = bool(checker(ring[0].numpy(), first))
            except Exception as exc:
                out["oracle_check_error"] = repr(exc)
    for lane in lanes:
        for b, _o in lane["batches"].values():
            b.close()
        lane["ctx"].close()
    return out
