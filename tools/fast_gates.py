"""Gate report of the tolerance mode of K1 (mnemo_create_ex(..., MNEMO_FAST_SPECTRA)) against the exact mode, on the
bench workloads: BASELINE.json configs[1] (one 2 h stereo track), configs[2] (130 four-minute tracks, a sample of them)
and configs[3] (10 000 thirty-second tracks, 10 000 five-second excerpts: the identified track must be the same for
every query).  Prints one JSON object; committed as profiles/r02_fast_gates.json.

    python tools/fast_gates.py [--tracks 10000]
"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mnemophonix_b200 import capi, synth
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--tracks", type=int, default=10000)
args = ap.parse_args()
dev = torch.device("cuda", 0)
exact, fast = capi.Context(0), capi.Context(0, fast_spectra=True)
out = {}


def compare(pcm, with_images):
    rec = {}
    if with_images:
        be, bf = exact.batch([pcm]), fast.batch([pcm])
        for b in (be, bf):
            b.enable_taps(True); b.upload(); b.run(); b.sync()
        img_e, img_f = be.read(6, 0).reshape(-1, 4096), bf.read(6, 0).reshape(-1, 4096)
        raw_e, raw_f = be.read(5, 0).reshape(-1, 1024), bf.read(5, 0).reshape(-1, 1024)
        rec["images"] = int(len(img_e))
        rec["image_max_abs_diff"] = float(np.abs(img_f - img_e).max())
        rec["raw_bit_agreement"] = float(1.0 - np.unpackbits(raw_e ^ raw_f).sum() / (raw_e.size * 8.0))
        same = (raw_e == raw_f).all(axis=1)
        rec["images_with_identical_raw_fingerprint"] = float(same.mean())
        se, sf = be.read(4, 0).reshape(-1, 100), bf.read(4, 0).reshape(-1, 100)
        rec["signatures_identical_where_raw_identical"] = bool(np.array_equal(se[same], sf[same]))
        be.close(); bf.close()
    se, _ = exact.index_pcm_batch([pcm]); sf, _ = fast.index_pcm_batch([pcm])
    rec["signatures"] = [int(len(se[0])), int(len(sf[0]))]
    if len(se[0]) == len(sf[0]):
        rec["signatures_identical_fraction"] = float((se[0] == sf[0]).all(axis=1).mean())
    return rec


# configs[1]: 10 minutes of the 2 h bench track with images (the taps of the whole track would need 2.5 GB x 3), then the whole track
out["c2_10min_stereo"] = compare(bench.make_track(600.0, 2, 1), True)
out["c2_2h_stereo"] = compare(bench.make_track(7200.0, 2, 1), False)
# configs[2]: 8 of the 130 four-minute mono tracks
c3 = [compare(synth.synth_track(1000 + i, 240.0, 1), True) for i in range(0, 130, 17)]
out["c3_sample"] = {"tracks": len(c3), "image_max_abs_diff": max(r["image_max_abs_diff"] for r in c3),
                    "raw_bit_agreement_min": min(r["raw_bit_agreement"] for r in c3),
                    "signatures_identical_where_raw_identical": all(r["signatures_identical_where_raw_identical"] for r in c3),
                    "signatures_identical_fraction_min": min(r.get("signatures_identical_fraction", 0.0) for r in c3)}
# configs[3]: the whole 10 k x 10 k search in both modes
q_off, q_len = int(10.0 * 44100) + 1234, 5 * 44100
answers, sig_sets = {}, {}
for name, ctx in (("exact", exact), ("fast", fast)):
    entries, queries = [], []
    for c0 in range(0, args.tracks, 50):
        ids = list(range(c0, min(args.tracks, c0 + 50)))
        pcm = synth.synth_tracks_torch(ids, 30.0, dev)
        host = torch.empty(pcm.shape, dtype=torch.int16).pin_memory(); host.copy_(pcm); torch.cuda.synchronize()
        arr = host.numpy()
        s, st = ctx.index_pcm_batch([arr[i].reshape(-1, 1) for i in range(len(ids))])
        q, qst = ctx.index_pcm_batch([np.ascontiguousarray(arr[i, q_off:q_off + q_len]).reshape(-1, 1) for i in range(len(ids))])
        entries += s; queries += q
    db = capi.Db(ctx, entries)
    res, _ = db.search(capi.Db.prepare(queries))
    db.close()
    answers[name] = [r[0] for r in res]
    sig_sets[name] = (entries, queries)
# and the mixed case: database indexed in exact mode, queries fingerprinted in tolerance mode
db = capi.Db(exact, sig_sets["exact"][0])
mixed = [r[0] for r in db.search(capi.Db.prepare(sig_sets["fast"][1]))[0]]
db.close()
n_same_sig = sum(int(np.array_equal(a, b)) for a, b in zip(sig_sets["exact"][0], sig_sets["fast"][0]))
out["c4"] = {"tracks": args.tracks, "queries": args.tracks,
             "identified_track_identical_on_every_query": answers["exact"] == answers["fast"],
             "queries_with_a_different_answer": int(sum(int(a != b) for a, b in zip(answers["exact"], answers["fast"]))),
             "own_track_identified": [int(sum(int(a == i) for i, a in enumerate(answers[m]))) for m in ("exact", "fast")],
             "exact_database_fast_queries_identical": mixed == answers["exact"],
             "tracks_with_byte_identical_signatures": n_same_sig}
out["gates_hold"] = bool(out["c2_10min_stereo"]["image_max_abs_diff"] <= 1e-4 and out["c3_sample"]["image_max_abs_diff"] <= 1e-4
                         and out["c2_10min_stereo"]["raw_bit_agreement"] >= 0.999 and out["c3_sample"]["raw_bit_agreement_min"] >= 0.999
                         and out["c2_10min_stereo"]["signatures_identical_where_raw_identical"]
                         and out["c3_sample"]["signatures_identical_where_raw_identical"]
                         and out["c4"]["identified_track_identical_on_every_query"])
print(json.dumps(out))
