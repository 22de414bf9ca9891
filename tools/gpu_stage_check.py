"""Stage-by-stage parity of the CUDA index path against the oracle / reference (dev tool)."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mnemophonix_b200 import capi, synth
from oracle.oracle import Oracle
from oracle import ref as refmod

o = Oracle()
r = refmod.Ref() if refmod.available() else None
ctx = capi.Context(0)
print("logf variant/probes/matched", ctx.logf_variant())

def eqbits(a, b):
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    if a.shape != b.shape: return False
    return np.array_equal(a.view(np.uint8), b.view(np.uint8))

ok_all = True
for seed, secs, ch in [(1, 30.0, 1), (2, 12.0, 2), (3, 3.0, 1), (4, 95.0, 2), (5, 20.0, 3)]:
    pcm = synth.synth_track(seed, secs, ch)
    b = ctx.batch([pcm]); b.enable_taps(True); b.upload(); b.run(); b.sync()
    s_gpu = b.read(0, 0); rms_gpu = b.read(1, 0)[0]; ss_gpu = b.read(8, 0)[0]
    s44 = o.downmix(pcm); s_or = o.resample(s44); n_or, rms_or = o.normalize(s_or)
    res = {"s5512": eqbits(s_gpu, s_or), "rms": np.float32(rms_gpu) == np.float32(rms_or)}
    rc, sig_or, taps = o.fingerprint_samples(n_or, taps=True)
    if rc == 0:
        bins_gpu = b.read(2, 0).reshape(-1, 32)
        # oracle bins: recompute per frame
        nf = bins_gpu.shape[0]
        bins_or = np.stack([o.frame_bins(n_or[64*f:64*f+2048]) for f in range(min(nf, 400))])
        res["bins"] = eqbits(bins_gpu[:len(bins_or)], bins_or)
        res["images"] = eqbits(b.read(6, 0).reshape(-1, 4096), taps["images"])
        res["haar"] = eqbits(b.read(7, 0).reshape(-1, 4096), taps["haar"])
        res["raw"] = eqbits(b.read(5, 0).reshape(-1, 1024), taps["raw"])
        sigs, st = b.signatures()
        res["sigs"] = eqbits(sigs[0], sig_or)
        res["n"] = (len(sigs[0]), len(sig_or))
        if not res["images"]:
            g = b.read(6, 0).reshape(-1, 4096); d = np.abs(g - taps["images"]); print("  img maxdiff", np.nanmax(d), "n diff", (g.view(np.uint32) != taps["images"].view(np.uint32)).sum())
        if not res["haar"]:
            g = b.read(7, 0).reshape(-1, 4096); print("  haar n diff", (g.view(np.uint32) != taps["haar"].view(np.uint32)).sum())
        if not res["bins"]:
            d = (bins_gpu[:len(bins_or)].view(np.uint32) != bins_or.view(np.uint32)); print("  bins n diff", d.sum(), "of", d.size, "first", np.argwhere(d)[:5])
    else:
        sigs, st = b.signatures(); res["status"] = (st, rc)
    print(seed, secs, ch, res)
    ok_all &= all(v for k, v in res.items() if isinstance(v, (bool, np.bool_)))
    b.close()
# from_samples path + tiny inputs
rc, sg = ctx.index_samples(n_or); print("index_samples", rc, eqbits(sg, sig_or))
rc, sg = ctx.index_samples(n_or[:5000]); print("index_samples small", rc)
sigs, st = ctx.index_pcm_batch([synth.synth_track(7, 0.5, 1), synth.synth_track(8, 8.0, 1), synth.synth_track(9, 1.0, 2)])
print("batch statuses", st, [len(s) for s in sigs])
rc8, s8 = o.fingerprint_pcm(synth.synth_track(8, 8.0, 1)); print("batch track1 eq", eqbits(sigs[1], s8))
t = time.time(); bad = ctx.selftest_logf(0x3f800000, 0x3f800000 + (1 << 22)); print("logf selftest mismatches (4M)", bad, time.time() - t)
print("ALL OK" if ok_all else "MISMATCH")
