"""Dev tool: where a config #4 search batch (10 k queries x 10 k tracks, signature-level synthetic) spends its time
on one GPU: result buffer, probe + push kernels, record download, exact host-side ranking."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mnemophonix_b200 import capi

tracks, S, nq = 10000, 304, 10000
def draw(rng, shape):
    return np.minimum(rng.geometric(1.0 / 38.0, size=shape) - 1, 254).astype(np.uint8)
rng = np.random.RandomState(7)
entries = draw(rng, (tracks, S, 100))
qtrack = rng.randint(0, tracks, nq); qoff = rng.randint(0, S - 34, nq)
queries = np.stack([entries[t, o:o + 34] for t, o in zip(qtrack, qoff)])
flip = rng.rand(*queries.shape) < 0.45
queries[flip] = draw(rng, int(flip.sum()))
prepared = capi.Db.prepare([queries[i] for i in range(nq)])
qstart = prepared[1]; nqs = int(qstart[-1])
stream = torch.cuda.Stream()
ctx = capi.Context(0, stream.cuda_stream)
t0 = time.perf_counter(); db = capi.Db(ctx, [entries[i] for i in range(tracks)]); print("db build s", time.perf_counter() - t0)
cap = 64 * nq
def T(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3, r
ms, (res, handle) = T(lambda: capi.Result.create(ctx, cap, 1, nqs)); print("Result.create  %.2f ms" % ms)
ms, _ = T(lambda: res.reset()); print("Result.reset   %.2f ms" % ms)
def push():
    res.reset(); return res.push(db, prepared, 0)
ms, st = T(push); print("reset+push     %.2f ms   L/qsig %.0f D/qsig %.1f" % (ms, st[0] / nqs, st[1] / nqs))
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
res.reset()
with torch.cuda.stream(stream):
    ev0.record(stream); res.push(db, prepared, 0); ev1.record(stream)
torch.cuda.synchronize(); print("push on-stream (H2D queries + kernels) %.2f ms" % ev0.elapsed_time(ev1))
ms, out = T(lambda: res.finish(qstart, tracks)); print("finish (D2H + host ranking) %.2f ms" % ms)
print("correct", sum(int(o[0] == t) for o, t in zip(out, qtrack)))
ms, (c, l, qs, stt) = T(lambda: db.search_shard(prepared)); print("search_shard (all-gather path, local part) %.2f ms, records %d" % (ms, len(c)))
ms, _ = T(lambda: capi.search_finish(c, None, 1, qstart, tracks)); print("search_finish %.2f ms" % ms)
ms, _ = T(lambda: db.search(prepared)); print("search_batch  %.2f ms -> %.0f q/s" % (ms, nq / ms * 1e3))
