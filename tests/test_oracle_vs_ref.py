"""CPU: the oracle against the live unmodified reference (oracle/_ref) on fresh inputs and edge cases."""
import ctypes as C
import os
import tempfile

import numpy as np
import pytest

from mnemophonix_b200 import synth


def bits_equal(a, b):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    return a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8))


@pytest.mark.parametrize("seed,secs,ch", [(21, 5.0, 1), (22, 7.3, 2), (23, 3.1, 4), (24, 2.0, 1)])
def test_every_stage_matches_reference(oracle, ref, seed, secs, ch):
    pcm = synth.synth_track(seed, secs, ch)
    s44 = oracle.downmix(pcm)
    assert bits_equal(s44, ref.downmix(pcm))
    s55 = oracle.resample(s44)
    assert bits_equal(s55, ref.resample(s44))
    norm, rms = oracle.normalize(s55)
    assert bits_equal(norm, ref.normalize(s55))
    st = ref.stages_from_samples(norm)
    rc, sigs, taps = oracle.fingerprint_samples(norm, taps=True)
    assert rc == st["rc"] == 0
    for k in ("images", "haar", "raw", "silence"):
        assert bits_equal(taps[k], st[k]), k
    assert bits_equal(sigs, st["sigs"])


def test_wav_file_path_matches(oracle, ref):
    pcm = synth.synth_track(31, 4.0, 2)
    p = tempfile.mktemp(suffix=".wav")
    synth.write_wav(p, pcm, {"IART": "artist", "INAM": "title", "IPRD": "album"})
    try:
        rc, sigs, a, t, al = ref.fingerprint_wav(p)
    finally:
        os.remove(p)
    assert rc == 0 and (a, t, al) == (b"artist", b"title", b"album")
    rc2, sigs2 = oracle.fingerprint_pcm(pcm)
    assert rc2 == 0 and bits_equal(sigs, sigs2)


def test_quiet_and_clipped_inputs(oracle, ref):
    rng = np.random.RandomState(3)
    quiet = (rng.standard_normal((44100 * 3, 1)) * 2).astype(np.int16)              # rms clamps to 0.1
    loud = (np.sign(rng.standard_normal((44100 * 3, 2))) * 32767).astype(np.int16)   # rms clamps to 3.0, clipping
    mixed = synth.synth_track(41, 4.0, 1)
    mixed[44100:2 * 44100] = 0                                                      # a silent second: silent images
    for pcm in (quiet, loud, mixed):
        rc_r, s_r = ref.fingerprint_pcm(pcm)
        rc_o, s_o = oracle.fingerprint_pcm(pcm)
        assert rc_r == rc_o
        if rc_r == 0:
            assert bits_equal(s_r, s_o)


def test_sequential_sum_of_squares_is_not_the_true_sum(oracle):
    # the quirk the GPU path must reproduce
    rng = np.random.RandomState(1)
    s = (rng.standard_normal(3_000_000) * 0.3).astype(np.float32)
    acc = np.float32(0)
    sq = s * s
    # float32 cumulative sum, sequential
    seq = np.cumsum(sq, dtype=np.float32)[-1]
    true = float(np.sum(sq.astype(np.float64)))
    norm, rms = oracle.normalize(s)
    expect = np.float32(np.float64(np.sqrt(seq / np.float32(len(s)))) * 10.0)
    assert np.float32(rms) == min(max(expect, np.float32(0.1)), np.float32(3.0))
    assert abs(float(seq) - true) / true > 1e-5


def test_search_matches_reference(oracle, ref):
    tracks = []
    for tid in range(4):
        rc, sg = oracle.fingerprint_pcm(synth.synth_track(3000 + tid, 12.0, 1))
        tracks.append(sg)
    p = tempfile.mktemp()
    with open(p, "w") as f:
        for i, sg in enumerate(tracks):
            f.write(f"t{i}\n\n\n\n{len(sg)}\n")
            for row in sg:
                f.write("".join(f"{b:02x}" for b in row) + "\n")
    rc, pidx = ref.load_index(p)
    os.remove(p)
    assert rc == 0
    lsh = ref.lib.create_hash_tables(pidx)
    db = oracle.make_db(tracks)
    for tid in range(4):
        pcm = synth.synth_track(3000 + tid, 12.0, 1)
        rc, q = oracle.fingerprint_pcm(pcm[3 * 44100 + 777: 8 * 44100 + 777])
        best, sc, L, D = oracle.search(db, q)
        assert best == ref.search(q, pidx, lsh) == tid
        assert oracle.get_matches(db, q[1]) == ref.get_matches(lsh, q[1])
    # self search exercises the dropped-last-group quirk (search.c:147-165)
    best, sc, L, D = oracle.search(db, tracks[3])
    assert best == ref.search(tracks[3], pidx, lsh)
    ref.lib.free_hash_tables(lsh)
    ref.lib.free_index(pidx)
