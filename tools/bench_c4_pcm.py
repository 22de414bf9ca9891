"""BASELINE config #4 from PCM: index M synthetic 30-second tracks, then search M five-second excerpts against
the resulting database, sharded across the GPUs of one box (contiguous entry ranges, records pushed into the
merging rank's HBM over NVLink, exact ranking there).

    python tools/bench_c4_pcm.py [--tracks 10000] [--seconds 30]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_c4_pcm.py

Unlike tools/bench_search_c4.py (signature-level random database: almost no bucket collisions, ~100 probed nodes per
query signature) the signatures here come from audio, so bucket occupancy is what a real database has (thousands of
probed nodes per query signature at this size).  10 k x 30 s of PCM (26 GB) cannot be synthesised by
the numpy generator in bench time, so the same recipe (repeating chirp + 3 voices of tone bursts + noise, SURVEY.md
§8d) is evaluated with torch on the GPU; the PCM then takes the normal route: device -> pinned host -> C ABI
(mnemo_batch_upload) -> kernels.  Rank 0 runs the unmodified reference (oracle/_ref; measurement tool, the reference
is the checker) on a sample of queries against the same database and compares the answers.
"""
import argparse, ctypes as C, json, math, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mnemophonix_b200 import capi, dist as mdist, synth

RATE = 44100
ap = argparse.ArgumentParser()
ap.add_argument("--tracks", type=int, default=10000)
ap.add_argument("--seconds", type=float, default=30.0)
ap.add_argument("--chunk", type=int, default=50)          # tracks synthesised + indexed per step
ap.add_argument("--ref-queries", type=int, default=8)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--cache", default="")                    # single GPU: keep / reuse the corpus signatures in this .npz
ap.add_argument("--l-hist", type=int, default=0)          # sample this many query signatures: distribution of L
ap.add_argument("--queries", type=int, default=0)         # 0 = one excerpt per track; else only the first N (profiling)
args = ap.parse_args()
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist = None
if world > 1:
    import torch.distributed as dist
    sys.stdout.flush(); saved = os.dup(1); os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=dev)
    dist.all_reduce(torch.zeros(1, device=dev)); torch.cuda.synchronize()
    sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)


def synth_gpu(track_ids, seconds):
    return synth.synth_tracks_torch(track_ids, seconds, dev)


# ---- this rank's contiguous range of tracks: synthesise, index, cut the query excerpts ----
cuts, _node_lo = capi.tree_cuts(args.tracks, world)      # shards on node boundaries of the ranking tree
a, b = int(cuts[rank]), int(cuts[rank + 1])
per = int(max(cuts[1:] - cuts[:-1]))
stream = torch.cuda.Stream()
ctx = capi.Context(lr, stream.cuda_stream)
n_frames = int(round(args.seconds * RATE))
q_off = int(10.0 * RATE) + 1234 if args.seconds < 70 else int(60.0 * RATE) + 1234
q_len = 5 * RATE
my_sigs, my_queries = [], []
t_synth = t_index = 0.0
cached = bool(args.cache) and world == 1 and os.path.exists(args.cache)
if cached:
    z = np.load(args.cache)
    zs, zst, zq, zqst = z["sigs"], z["start"], z["qsigs"], z["qstart"]     # (an NpzFile re-reads the array on every access)
    my_sigs = [zs[zst[i]:zst[i + 1]] for i in range(len(zst) - 1)]
    my_queries = [zq[zqst[i]:zqst[i + 1]] for i in range(len(zqst) - 1)]
for c0 in ([] if cached else range(a, b, args.chunk)):
    ids = list(range(c0, min(b, c0 + args.chunk)))
    t0 = time.perf_counter()
    pcm = synth_gpu(ids, args.seconds)
    host = torch.empty(pcm.shape, dtype=torch.int16).pin_memory()
    host.copy_(pcm); torch.cuda.synchronize()
    t_synth += time.perf_counter() - t0
    t0 = time.perf_counter()
    arr = host.numpy()
    tracks = [arr[i].reshape(-1, 1) for i in range(len(ids))]
    sigs, status = ctx.index_pcm_batch(tracks)
    clips = [np.ascontiguousarray(arr[i, q_off:q_off + q_len]).reshape(-1, 1) for i in range(len(ids))]
    qs, qstatus = ctx.index_pcm_batch(clips)
    assert not any(status) and not any(qstatus)
    my_sigs += sigs; my_queries += qs
    t_index += time.perf_counter() - t0
    del pcm, host

if args.cache and world == 1 and not cached:
    np.savez(args.cache, sigs=np.concatenate(my_sigs), start=np.cumsum([0] + [len(x) for x in my_sigs]),
             qsigs=np.concatenate(my_queries), qstart=np.cumsum([0] + [len(x) for x in my_queries]))
counts = torch.zeros(args.tracks, dtype=torch.int64, device=dev)
counts[a:b] = torch.tensor([len(s) for s in my_sigs], dtype=torch.int64)
if dist is not None:
    dist.all_reduce(counts)
total_sigs = int(counts.sum())

# queries: every rank needs all of them (replicated to all shards): all-gather the signatures
qcnt = torch.zeros(args.tracks, dtype=torch.int64, device=dev)
qcnt[a:b] = torch.tensor([len(q) for q in my_queries], dtype=torch.int64)
if dist is not None:
    dist.all_reduce(qcnt)
qcnt = qcnt.cpu().numpy()
qmax = int(qcnt.max())
mine = torch.zeros((per, qmax, 100), dtype=torch.uint8)
for i, q in enumerate(my_queries):
    mine[i, :len(q)] = torch.from_numpy(np.ascontiguousarray(q))
if dist is not None:
    allq = [torch.empty((per, qmax, 100), dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(allq, mine.to(dev))
    allq = torch.cat([allq[r][:int(cuts[r + 1] - cuts[r])] for r in range(world)]).cpu().numpy()
else:
    allq = mine.numpy()
n_q = args.queries or args.tracks
queries = [allq[i, :qcnt[i]] for i in range(n_q)]
prepared = capi.Db.prepare(queries)
n_qsig = int(prepared[1][-1])

t0 = time.perf_counter()
searcher = mdist.RankedShardedSearcher(ctx, my_sigs, total_sigs // 2, args.tracks, rank, world, device=dev)
torch.cuda.synchronize()
t_build = time.perf_counter() - t0


def timed(fn, steps):
    ts, out = [], None
    for _ in range(steps):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if dist is not None:
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        ts.append(dt)
    return float(np.mean(ts)), out


if world == 1:
    # one shard: probe + exact ranking on the GPU (mnemo_search_batch), 12 bytes per query come back
    searcher.db.search(prepared)
    dt, res = timed(lambda: searcher.db.search(prepared)[0], args.steps)
else:
    # shards ranked on their GPUs, two small all-gathers, host merge of ten records per query and node
    searcher.search(prepared)
    dt, res = timed(lambda: searcher.search(prepared), args.steps)
L = D = 0
if world == 1:
    L, D = searcher.db.search(prepared)[1]

breakdown = None
l_hist = None
if args.l_hist and world == 1:
    rs = np.random.RandomState(5)
    pick_q = rs.randint(0, n_qsig, args.l_hist)
    one = np.empty((1, 2), np.uint32)
    Ls = np.array([capi.lib().mnemo_db_get_matches(searcher.db.h, np.ascontiguousarray(prepared[0][i]).ctypes.data,
                                                   one.ctypes.data, 1) for i in pick_q], dtype=np.int64)
    # time the heaviest and the lightest sampled signatures as single-signature queries
    order = np.argsort(Ls)
    def t_one(ix):
        q = [prepared[0][pick_q[ix]][None, :]]
        searcher.db.search(q); torch.cuda.synchronize(); t0 = time.perf_counter(); searcher.db.search(q); torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3
    l_hist = {"n": int(len(Ls)), "mean": float(Ls.mean()), "p50": float(np.percentile(Ls, 50)), "p90": float(np.percentile(Ls, 90)),
              "p99": float(np.percentile(Ls, 99)), "max": int(Ls.max()), "share_of_nodes_in_top_1pct": float(np.sort(Ls)[-max(1, len(Ls) // 100):].sum() / Ls.sum()),
              "ms_single_sig_query_lightest": t_one(order[0]), "ms_single_sig_query_median": t_one(order[len(order) // 2]),
              "ms_single_sig_query_p99": t_one(order[int(len(order) * 0.99)]), "ms_single_sig_query_heaviest": t_one(order[-1]),
              "L_lightest": int(Ls[order[0]]), "L_median": int(Ls[order[len(order) // 2]]), "L_p99": int(Ls[order[int(len(order) * 0.99)]])}
if rank == 0:
    got = [r[0] for r in res]
    line = {"metric": "search queries/s",
            "config": "BASELINE configs[3] from PCM: %d five-second excerpts x %d-track database (%g s tracks, %d signatures from "
                      "GPU-synthesised audio), %d GPU(s), shards ranked on their GPUs" % (args.tracks, args.tracks, args.seconds, total_sigs, world),
            "n_gpus": world, "queries": n_q, "query_signatures": n_qsig, "queries_per_s": n_q / dt,
            "s_per_batch": dt, "identified_own_track": int(sum(int(g == i) for i, g in enumerate(got))),
            "no_match": int(sum(int(g < 0) for g in got)), "lsh_build_s_per_shard": t_build,
            "index_s_this_rank": t_index, "synth_s_this_rank": t_synth}
    if world == 1:
        line.update({"L_distribution": l_hist, "probed_nodes_per_query_signature": L / n_qsig, "deep_checks_per_query_signature": D / n_qsig,
                     "algorithmic_GBps": (n_qsig * 300 + 4 * L + 100 * D + 16 * n_q) / dt / 1e9})
    print(json.dumps(line)); sys.stdout.flush()

# ---- reference on a sample (rank 0 needs the whole database for that) ----
if args.ref_queries:
    smax = int(counts.max())
    mine_s = torch.zeros((per, smax, 100), dtype=torch.uint8)
    for i, s in enumerate(my_sigs):
        mine_s[i, :len(s)] = torch.from_numpy(np.ascontiguousarray(s))
    if dist is not None:
        alls = [torch.empty((per, smax, 100), dtype=torch.uint8, device=dev) for _ in range(world)]
        dist.all_gather(alls, mine_s.to(dev))
        alls = torch.cat([alls[r][:int(cuts[r + 1] - cuts[r])] for r in range(world)]).cpu().numpy()
    else:
        alls = mine_s.numpy()
    if rank == 0:
        from oracle import ref as refmod
        if refmod.available():
            cn = counts.cpu().numpy()
            r = refmod.Ref()
            ents = (C.POINTER(refmod.IndexEntry) * args.tracks)()
            keep = []
            for i in range(args.tracks):
                e = np.ascontiguousarray(alls[i, :cn[i]])
                sg = refmod.Signatures(int(cn[i]), C.cast(e.ctypes.data, C.POINTER(C.c_uint8 * 100)))
                ie = refmod.IndexEntry(b"t", b"", b"", b"", C.pointer(sg))
                keep.append((e, sg, ie)); ents[i] = C.pointer(ie)
            idx = refmod.Index(args.tracks, C.cast(ents, C.POINTER(C.POINTER(refmod.IndexEntry))))
            t0 = time.perf_counter(); lsh = r.lib.create_hash_tables(C.byref(idx)); t_lsh = time.perf_counter() - t0
            pick = np.linspace(0, n_q - 1, args.ref_queries).astype(int)
            t0 = time.perf_counter()
            ref_res = [r.search(np.ascontiguousarray(queries[i]), C.pointer(idx), lsh) for i in pick]
            t_ref = (time.perf_counter() - t0) / len(pick)
            print(json.dumps({"ref_create_hash_tables_s": t_lsh, "ref_search_s_per_query": t_ref, "ref_sample": len(pick),
                              "identical_results_on_sample": ref_res == [got[i] for i in pick]}))
searcher.close(); ctx.close()
if dist is not None:
    dist.destroy_process_group()
