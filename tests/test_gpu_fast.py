"""GPU: the TOLERANCE mode of the frame -> band-energy kernel (mnemo_create_ex(..., MNEMO_FAST_SPECTRA), k1_spectral.cu
"fast K1") against the exact mode, with BASELINE.json's tolerances written out:
  * spectral images within 1e-4 (they are scaled to [0, 1], spectralimages.c:52-77; relative to that full scale);
  * raw fingerprint bits agree on >= 99.9 %;
  * signatures are IDENTICAL for every image whose raw fingerprint is identical (everything after K1 is the same
    exact code);
  * the search identifies the same track for every excerpt.
The default mode stays the bit-exact one: every other GPU test runs on it."""
import numpy as np
import pytest

from mnemophonix_b200 import capi, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fast():
    c = capi.Context(0, fast_spectra=True)
    yield c
    c.close()


def _taps(ctx, pcm):
    b = ctx.batch([pcm])
    b.enable_taps(True)
    b.upload()
    b.run()
    b.sync()
    return b


def test_modes_are_what_they_say(ctx, fast):
    assert capi.lib().mnemo_spectra_mode(ctx.h) == 0 and capi.lib().mnemo_spectra_mode(fast.h) == 1


@pytest.mark.parametrize("seed,seconds,channels", [(21, 20.0, 1), (22, 33.0, 2), (23, 12.5, 1)])
def test_fast_spectra_within_the_contract_tolerances(ctx, fast, seed, seconds, channels):
    pcm = synth.synth_track(seed, seconds, channels)
    be, bf = _taps(ctx, pcm), _taps(fast, pcm)
    bins_e, bins_f = be.read(2, 0).astype(np.float64), bf.read(2, 0).astype(np.float64)
    frame_max = bins_e.reshape(-1, 32).max(axis=1, keepdims=True)
    rel = np.abs(bins_f - bins_e).reshape(-1, 32) / np.maximum(frame_max, 1e-30)
    assert rel.max() < 2e-5, rel.max()                                  # band energies: ~1e-6 typical
    img_e, img_f = be.read(6, 0).reshape(-1, 4096), bf.read(6, 0).reshape(-1, 4096)
    assert np.abs(img_f - img_e).max() <= 1e-4                          # BASELINE.json: images within 1e-4
    raw_e, raw_f = be.read(5, 0).reshape(-1, 1024), bf.read(5, 0).reshape(-1, 1024)
    differing = np.unpackbits(raw_e ^ raw_f).sum()
    assert differing <= 1e-3 * raw_e.size * 8, differing               # raw bits agree on >= 99.9 %
    same = (raw_e == raw_f).all(axis=1)
    sig_e, sig_f = be.read(4, 0).reshape(-1, 100), bf.read(4, 0).reshape(-1, 100)
    keep_e, keep_f = be.read(3, 0), bf.read(3, 0)
    assert np.array_equal(sig_e[same], sig_f[same]) and np.array_equal(keep_e[same], keep_f[same])
    assert same.mean() > 0.9
    be.close()
    bf.close()


def test_fast_spectra_identify_the_same_tracks(ctx, fast):
    pcms = [synth.synth_track(4100 + i, 22.0 + 2 * (i % 3), 1 + (i % 2)) for i in range(10)]
    clips = [np.ascontiguousarray(p[5 * 44100 + 777: 10 * 44100 + 777]) for p in pcms]
    answers = []
    for c in (ctx, fast):
        sigs, st = c.index_pcm_batch(pcms)
        qs, _ = c.index_pcm_batch(clips)
        db = capi.Db(c, sigs)
        res, _ = db.search(qs)
        db.close()
        answers.append([r[0] for r in res])
    assert answers[0] == answers[1] == list(range(10))
    # a database built in one mode answers queries fingerprinted in the other
    sigs_e, _ = ctx.index_pcm_batch(pcms)
    qs_f, _ = fast.index_pcm_batch(clips)
    db = capi.Db(ctx, sigs_e)
    assert [r[0] for r in db.search(qs_f)[0]] == list(range(10))
    db.close()
