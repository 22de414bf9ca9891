"""Sparse-regime search case: N query clips against an M-track database that is synthetic at the SIGNATURE level
(bytes follow the geometric-like distribution measured on real signatures; queries are 34-signature
excerpts with ~45 % of the bytes re-drawn).  Such a database has almost no bucket collisions (L ~ 100 probed nodes per
query signature), so this exercises the probe kernel's hashed mode; tools/bench_c4_pcm.py is BASELINE config #4 proper
(signatures from audio, L ~ 42 k).  One GPU.

    python tools/bench_search_c4.py [--tracks 10000] [--sigs-per-track 304] [--queries 10000]
"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mnemophonix_b200 import capi

ap = argparse.ArgumentParser()
ap.add_argument("--tracks", type=int, default=10000)
ap.add_argument("--sigs-per-track", type=int, default=304)
ap.add_argument("--queries", type=int, default=10000)
ap.add_argument("--steps", type=int, default=3)
args = ap.parse_args()


def draw(rng, shape):
    return np.minimum(rng.geometric(1.0 / 38.0, size=shape) - 1, 254).astype(np.uint8)


rng = np.random.RandomState(7)
S = args.sigs_per_track
entries = draw(rng, (args.tracks, S, 100))
qtrack = rng.randint(0, args.tracks, args.queries)
qoff = rng.randint(0, S - 34, args.queries)
queries = np.stack([entries[t, o:o + 34] for t, o in zip(qtrack, qoff)])
flip = rng.rand(*queries.shape) < 0.45
queries[flip] = draw(rng, int(flip.sum()))
prepared = capi.Db.prepare([queries[i] for i in range(args.queries)])

stream = torch.cuda.Stream()
ctx = capi.Context(0, stream.cuda_stream)
t0 = time.perf_counter()
db = capi.Db(ctx, [entries[i] for i in range(args.tracks)])
t_build = time.perf_counter() - t0
db.search(prepared)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(args.steps):
    res, (L, D) = db.search(prepared)
dt = (time.perf_counter() - t0) / args.steps
nqs = int(prepared[1][-1])
print(json.dumps({"metric": "search queries/s", "config": "sparse regime: %d query clips x %d-track signature-level random database "
                  "(%d signatures), one GPU" % (args.queries, args.tracks, args.tracks * S), "queries_per_s": args.queries / dt,
                  "s_per_batch": dt, "lsh_build_s": t_build, "probed_nodes_per_query_signature": L / nqs,
                  "deep_checks_per_query_signature": D / nqs,
                  "identified_correct_track": int(sum(int(r[0] == t) for r, t in zip(res, qtrack)))}))
db.close(); ctx.close()
