"""Dev tool: time the single-GPU config #4 search (10 k excerpts x 10 k-track database from GPU-synthesised PCM) with
whatever libmnemo_cuda.so is in place; the corpus is cached in /tmp so that several library variants can be timed in
one gpurun call (tools/build_variants.sh).  Prints one JSON line: ms per batch, answers CRC (must not change).

    python tools/pair_ab.py [label] [tracks]      # tracks < 10000: a shard (first tracks, the whole database's table size)"""
import json, os, sys, time, zlib
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mnemophonix_b200 import capi, synth

label = sys.argv[1] if len(sys.argv) > 1 else ""
n_shard = int(sys.argv[2]) if len(sys.argv) > 2 else 0
cache = "/tmp/c4_corpus.npz"
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream()
ctx = capi.Context(0, stream.cuda_stream)
if os.path.exists(cache):
    z = np.load(cache)
    sig_all, sig_cnt, q_all, q_cnt = z["sig_all"], z["sig_cnt"], z["q_all"], z["q_cnt"]
else:
    n_tracks, seconds, chunk = 10000, 30.0, 50
    q_off, q_len = int(10.0 * 44100) + 1234, 5 * 44100
    sigs, qs = [], []
    for c0 in range(0, n_tracks, chunk):
        ids = list(range(c0, c0 + chunk))
        pcm = synth.synth_tracks_torch(ids, seconds, dev)
        host = torch.empty(pcm.shape, dtype=torch.int16).pin_memory()
        host.copy_(pcm)
        torch.cuda.synchronize()
        arr = host.numpy()
        s, st = ctx.index_pcm_batch([arr[i].reshape(-1, 1) for i in range(len(ids))])
        q, qst = ctx.index_pcm_batch([np.ascontiguousarray(arr[i, q_off:q_off + q_len]).reshape(-1, 1) for i in range(len(ids))])
        sigs += s
        qs += q
    sig_all, sig_cnt = np.concatenate(sigs), np.array([len(x) for x in sigs])
    q_all, q_cnt = np.concatenate(qs), np.array([len(x) for x in qs])
    np.savez(cache, sig_all=sig_all, sig_cnt=sig_cnt, q_all=q_all, q_cnt=q_cnt)
entries = np.split(sig_all, np.cumsum(sig_cnt)[:-1])
queries = np.split(q_all, np.cumsum(q_cnt)[:-1])
if n_shard:
    db = capi.Db(ctx, entries[:n_shard], table_size=len(sig_all) // 2)
else:
    db = capi.Db(ctx, entries)
prepared = capi.Db.prepare(queries, ctx)
db.search(prepared)
db.search(prepared)
times = []
for _ in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res, (L, D) = db.search(prepared)
    times.append((time.perf_counter() - t0) * 1e3)
crc = zlib.crc32(np.array(res, dtype=np.float64).tobytes())
print(json.dumps({"label": label, "tracks": n_shard or len(entries), "pair_index": db.has_pair_index, "ms_per_batch": float(np.median(times)), "ms_each": times, "queries_per_s": len(queries) / (np.median(times) / 1e3),
                  "answers_crc": crc, "L": L, "D": D, "identified": int(sum(int(r[0] == i) for i, r in enumerate(res)))}))
db.close()
ctx.close()
