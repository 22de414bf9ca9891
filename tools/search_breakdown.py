"""Dev tool: where a search batch spends its time (shard kernel + copies vs host-side finish)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mnemophonix_b200 import capi, synth
ctx = capi.Context(0)
pcm = np.ascontiguousarray(np.tile(synth.synth_track(1, 600.0, 2), (12, 1)))
sigs, st = ctx.index_pcm_batch([pcm])
sigs = sigs[0]
per = 304
entries = [sigs[i:i + per] for i in range(0, len(sigs) - per + 1, per)]
rng = np.random.RandomState(3)
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
picks = rng.randint(0, len(entries), nq); offs = rng.randint(0, per - 34, nq)
queries = [entries[e][o:o + 34] for e, o in zip(picks, offs)]
db = capi.Db(ctx, entries)
prep = capi.Db.prepare(queries)
db.search(prep)
def t(fn, reps=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps): r = fn()
    return (time.perf_counter() - t0) / reps, r
dt, (res, (L, D)) = t(lambda: db.search(prep))
print("search_batch (python call)  %.2f ms  -> %.0f q/s" % (dt * 1e3, nq / dt))
dt2, (c, l, qstart, stats) = t(lambda: db.search_shard(prep, cap=nq * 400))
print("search_shard, ample cap     %.2f ms   records %d (%.1f per query)" % (dt2 * 1e3, len(c), len(c) / nq))
dt3, out = t(lambda: capi.search_finish(c, None, 1, qstart, len(entries)))
print("search_finish               %.2f ms" % (dt3 * 1e3))
