"""Small end-to-end run for compute-sanitizer (one tool per gpurun call):
    compute-sanitizer --tool memcheck  python tools/gpu_sanitize.py
    compute-sanitizer --tool racecheck python tools/gpu_sanitize.py
Covers every kernel: index (ragged batch, too-small track, silence), from_samples, LSH build, search,
sharded search records, get_matches.  Results are still checked against the oracle."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mnemophonix_b200 import capi, synth
from oracle.oracle import Oracle

o = Oracle()
ctx = capi.Context(0)
pcms = [synth.synth_track(1, 4.5, 2), synth.synth_track(2, 2.2, 1), synth.synth_track(3, 0.3, 1), np.zeros((90000, 1), np.int16),
        synth.synth_track(4, 3.0, 3)]
got, st = ctx.index_pcm_batch(pcms)
for pcm, g, s in zip(pcms, got, st):
    rc, sigs = o.fingerprint_pcm(pcm)
    assert s == rc and np.array_equal(g, sigs), "index mismatch"
s = np.clip(np.random.RandomState(0).standard_normal(12000) * 0.3, -1, 1).astype(np.float32)
rc, g = ctx.index_samples(s)
assert rc == 0 and np.array_equal(g, o.fingerprint_samples(s)[1])
entries = [got[0], got[1], got[4]]
db = capi.Db(ctx, entries)
odb = o.make_db(entries)
queries = [got[0][3:20], got[4][:10], np.random.RandomState(1).randint(0, 255, (9, 100)).astype(np.uint8)]
res, (L, D) = db.search(queries)
for q, r in zip(queries, res):
    assert r[0] == o.search(odb, q)[0], "search mismatch"
c, l, qs, _ = db.search_shard(queries)
assert db.get_matches(got[0][5]) == o.get_matches(odb, got[0][5])
db.close(); ctx.close()
print("sanitize run OK: index/search results identical to the oracle")
