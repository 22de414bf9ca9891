"""BASELINE config #4: N query clips against an M-track database sharded across the GPUs of one box
(contiguous entry ranges, one all-gather of candidate records over NCCL, exact ranking on the host).

    python tools/bench_search_c4.py [--tracks 10000] [--sigs-per-track 304] [--queries 10000]
    torchrun --nproc-per-node 8 tools/bench_search_c4.py ...

The database is synthetic at the SIGNATURE level (10 k x 30 s of PCM cannot be synthesised on the host in
bench time): bytes follow the geometric-like distribution measured on real signatures (SURVEY.md §7 H6,
mean ~37), queries are 34-signature excerpts with ~45 % of the bytes re-drawn.  Rank 0 also runs the
unmodified reference (oracle/_ref, in-memory struct index) on a sample of queries and checks the answers."""
import argparse, ctypes as C, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mnemophonix_b200 import capi, dist as mdist
from oracle import ref as refmod

ap = argparse.ArgumentParser()
ap.add_argument("--tracks", type=int, default=10000)
ap.add_argument("--sigs-per-track", type=int, default=304)
ap.add_argument("--queries", type=int, default=10000)
ap.add_argument("--ref-queries", type=int, default=8)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--merge", default="both", choices=["allgather", "push", "both"])
args = ap.parse_args()
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))

def draw(rng, shape):
    return np.minimum(rng.geometric(1.0 / 38.0, size=shape) - 1, 254).astype(np.uint8)

# every rank generates the same corpus deterministically (a real deployment would load its shard)
rng = np.random.RandomState(7)
S = args.sigs_per_track
entries = draw(rng, (args.tracks, S, 100))
qtrack = rng.randint(0, args.tracks, args.queries)
qoff = rng.randint(0, S - 34, args.queries)
queries = np.stack([entries[t, o:o + 34] for t, o in zip(qtrack, qoff)])
flip = rng.rand(*queries.shape) < 0.45
queries[flip] = draw(rng, int(flip.sum()))
entry_list = [entries[i] for i in range(args.tracks)]
query_list = [queries[i] for i in range(args.queries)]
prepared = capi.Db.prepare(query_list)   # input marshalling is not part of the search

stream = torch.cuda.Stream()
ctx = capi.Context(lr, stream.cuda_stream)
dev = torch.device("cuda", lr)
t0 = time.perf_counter()
res, db = mdist.search_sharded(ctx, entry_list, query_list[:64], rank, world, dev)     # builds the shard + warm-up
torch.cuda.synchronize()
t_build = time.perf_counter() - t0
def timed(fn):
    ts, out = [], None
    for _ in range(args.steps):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        ts.append(dt)
    return float(np.mean(ts)), out

times_push = None
if args.merge in ("push", "both"):
    mdist.search_sharded_push(ctx, entry_list, prepared, rank, world, db=db)          # warm-up (maps the buffer once)
    times_push, res_push = timed(lambda: mdist.search_sharded_push(ctx, entry_list, prepared, rank, world, db=db)[0])
if args.merge in ("allgather", "both"):
    t_ag, res = timed(lambda: mdist.search_sharded(ctx, entry_list, prepared, rank, world, dev, db=db)[0])
    times = [t_ag]
else:
    res, times = res_push, [times_push]
if rank == 0:
    got = [r[0] for r in res]
    correct = int(sum(int(g == t) for g, t in zip(got, qtrack)))
    line = {"metric": "search queries/s", "config": "BASELINE configs[3]: %d query clips x %d-track DB (%d signatures), %d GPU(s)" %
            (args.queries, args.tracks, args.tracks * S, world), "n_gpus": world, "queries_per_s": args.queries / float(np.mean(times)),
            "s_per_batch": float(np.mean(times)), "shard_build_and_warmup_s": t_build, "identified_correct_track": correct,
            "queries": args.queries, "merge": "NCCL all-gather of records" if args.merge != "push" else "peer-memory push"}
    if times_push is not None and args.merge == "both":
        line["peer_push_queries_per_s"] = args.queries / times_push
        line["peer_push_s_per_batch"] = times_push
        line["peer_push_equals_allgather"] = res_push == res
    if refmod.available() and args.ref_queries:
        r = refmod.Ref()
        # in-memory struct index for the reference (no 600 MB text file)
        ents = (C.POINTER(refmod.IndexEntry) * args.tracks)()
        keep = []
        for i in range(args.tracks):
            sg = refmod.Signatures(S, C.cast(entries[i].ctypes.data, C.POINTER(C.c_uint8 * 100)))
            e = refmod.IndexEntry(b"t", b"", b"", b"", C.pointer(sg))
            keep.append((sg, e))
            ents[i] = C.pointer(e)
        idx = refmod.Index(args.tracks, C.cast(ents, C.POINTER(C.POINTER(refmod.IndexEntry))))
        t0 = time.perf_counter(); lsh = r.lib.create_hash_tables(C.byref(idx)); t_lsh = time.perf_counter() - t0
        t0 = time.perf_counter()
        ref_res = [r.search(queries[i], C.pointer(idx), lsh) for i in range(args.ref_queries)]
        t_ref = (time.perf_counter() - t0) / args.ref_queries
        line.update({"ref_create_hash_tables_s": t_lsh, "ref_search_s_per_query": t_ref, "ref_sample": args.ref_queries,
                     "identical_results_on_sample": ref_res == got[:args.ref_queries]})
    print(json.dumps(line))
db.close(); ctx.close()
if dist is not None:
    dist.destroy_process_group()
