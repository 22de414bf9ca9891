"""Search benchmark (BASELINE.json config #3 scale: 130 x 240 s tracks, 5 s query clips).

    python tools/bench_search.py [--tracks 130] [--seconds 240] [--queries 130] [--ref-queries 16]

Indexes the synthetic corpus on the GPU, builds the LSH tables on the GPU, searches all query clips in one
batch, and (when oracle/_ref is present) times the unmodified reference's single-threaded search() on a
sample of the same queries, checking that the identified entries agree.  Prints one JSON line."""
import argparse, ctypes as C, json, os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mnemophonix_b200 import capi, synth
from oracle import ref as refmod

ap = argparse.ArgumentParser()
ap.add_argument("--tracks", type=int, default=130)
ap.add_argument("--seconds", type=float, default=240.0)
ap.add_argument("--queries", type=int, default=130)
ap.add_argument("--ref-queries", type=int, default=16)
ap.add_argument("--steps", type=int, default=5)
args = ap.parse_args()

stream = torch.cuda.Stream()
ctx = capi.Context(0, stream.cuda_stream)
t0 = time.perf_counter()
sigs, clips = [], []
for base in range(0, args.tracks, 26):
    pcms = [synth.synth_track(1000 + i, args.seconds, 1) for i in range(base, min(args.tracks, base + 26))]
    s, st = ctx.index_pcm_batch(pcms)
    sigs += s
    off = int(min(60.0, args.seconds / 3) * 44100) + 1234
    clips += [np.ascontiguousarray(p[off: off + 5 * 44100]) for p in pcms]
t_index = time.perf_counter() - t0
qs, _ = ctx.index_pcm_batch(clips[:args.queries])
total = sum(len(s) for s in sigs)

t0 = time.perf_counter()
db = capi.Db(ctx, sigs)
t_build = time.perf_counter() - t0
db.search(qs[:4])                                   # warm-up
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
ev0.record(stream)
for _ in range(args.steps):
    res, (L, D) = db.search(qs)
ev1.record(stream)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / args.steps
correct = sum(r[0] == i for i, r in enumerate(res))
n_qsig = sum(len(q) for q in qs)
line = {"metric": "search queries/s", "db_tracks": args.tracks, "db_signatures": total, "queries": len(qs),
        "query_signatures": n_qsig, "gpu_db_build_s": t_build, "gpu_search_s_per_batch": wall,
        "gpu_queries_per_s": len(qs) / wall, "correct": correct,
        "L_per_query_signature": L / n_qsig, "D_per_query_signature": D / n_qsig,
        "algorithmic_bytes_per_batch": n_qsig * (100 + 25 * 8) + 4 * L + 100 * D + 16 * len(qs),
        "index_corpus_s_incl_synthesis": t_index}
line["algorithmic_GBps"] = line["algorithmic_bytes_per_batch"] / wall / 1e9

if refmod.available():
    r = refmod.Ref()
    p = tempfile.mktemp()
    with open(p, "w") as f:
        for i, sg in enumerate(sigs):
            f.write(f"t{i}\n\n\n\n{len(sg)}\n")
            f.write("".join("".join(f"{b:02x}" for b in row) + "\n" for row in sg))
    t0 = time.perf_counter(); rc, pidx = r.load_index(p); t_load = time.perf_counter() - t0
    os.remove(p)
    t0 = time.perf_counter(); lsh = r.lib.create_hash_tables(pidx); t_lsh = time.perf_counter() - t0
    nref = min(args.ref_queries, len(qs))
    t0 = time.perf_counter()
    ref_res = [r.search(qs[i], pidx, lsh) for i in range(nref)]
    t_ref = (time.perf_counter() - t0) / nref
    line.update({"ref_read_index_s": t_load, "ref_create_hash_tables_s": t_lsh, "ref_search_s_per_query": t_ref,
                 "ref_queries_per_s_1thread": 1.0 / t_ref, "ref_sample": nref,
                 "identical_results_on_sample": ref_res == [x[0] for x in res[:nref]]})
print(json.dumps(line))
