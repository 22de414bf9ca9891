"""GPU: parity of the CUDA index path (through the C ABI) with the oracle, the golden vectors and,
when oracle/_ref is present, the unmodified reference.  Bar: bit-exact at every stage."""
import numpy as np
import pytest

from mnemophonix_b200 import synth

pytestmark = pytest.mark.gpu


def bits_equal(a, b):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    return a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8))


def run_with_taps(ctx, pcm):
    b = ctx.batch([pcm])
    b.enable_taps(True)
    b.upload()
    b.run()
    b.sync()
    return b


@pytest.mark.parametrize("seed,secs,ch", [(51, 30.0, 1), (52, 9.7, 2), (53, 2.4, 1), (54, 61.0, 2), (55, 6.0, 5)])
def test_every_stage_is_bit_exact(ctx, oracle, seed, secs, ch):
    pcm = synth.synth_track(seed, secs, ch)
    b = run_with_taps(ctx, pcm)
    s55 = oracle.resample(oracle.downmix(pcm))
    assert bits_equal(b.read(0, 0), s55)                                   # wav.c:358-375 + resample.c
    norm, rms = oracle.normalize(s55)
    assert np.float32(b.read(1, 0)[0]) == np.float32(rms)                  # audionormalizer.c:5-20
    rc, sigs, taps = oracle.fingerprint_samples(norm, taps=True)
    assert rc == 0
    nf = len(b.read(2, 0)) // 32
    step = max(1, nf // 150)
    frames = list(range(0, nf, step)) + [nf - 1]
    bins_gpu = b.read(2, 0).reshape(-1, 32)
    bins_or = np.stack([oracle.frame_bins(norm[64 * f:64 * f + 2048]) for f in frames])
    assert bits_equal(bins_gpu[frames], bins_or)                           # spectralimages.c:80-107
    assert bits_equal(b.read(6, 0).reshape(-1, 4096), taps["images"])      # spectralimages.c:52-77
    assert bits_equal(b.read(7, 0).reshape(-1, 4096), taps["haar"])        # haar.c:23-73
    assert bits_equal(b.read(5, 0).reshape(-1, 1024), taps["raw"])         # rawfingerprints.c:43-100
    got, st = b.signatures()
    assert st == [0] and bits_equal(got[0], sigs)                          # minhash.c:13-54
    b.close()


def test_golden_vectors(ctx, golden):
    for name in ("a", "b"):
        pcm = np.ascontiguousarray(golden[f"{name}_pcm"])
        b = run_with_taps(ctx, pcm)
        assert bits_equal(b.read(0, 0), golden[f"{name}_s5512"])
        n = len(golden[f"{name}_bins_head"])
        assert bits_equal(b.read(2, 0).reshape(-1, 32)[:n], golden[f"{name}_bins_head"])
        pick = golden[f"{name}_pick"]
        assert bits_equal(b.read(6, 0).reshape(-1, 4096)[pick], golden[f"{name}_images"])
        assert bits_equal(b.read(7, 0).reshape(-1, 4096)[pick], golden[f"{name}_haar"])
        assert bits_equal(b.read(5, 0).reshape(-1, 1024), golden[f"{name}_raw"])
        got, st = b.signatures()
        assert bits_equal(got[0], golden[f"{name}_sigs"])
        b.close()


def test_against_live_reference(ctx, ref):
    pcm = synth.synth_track(61, 14.0, 2)
    rc, sigs_ref = ref.fingerprint_pcm(pcm)
    got, st = ctx.index_pcm_batch([pcm])
    assert rc == 0 and st == [0] and bits_equal(got[0], sigs_ref)


def test_ragged_batch_and_too_small_tracks(ctx, oracle):
    pcms = [synth.synth_track(71, 0.4, 1),            # < 2048 samples at 5512 Hz
            synth.synth_track(72, 5.0, 2),
            synth.synth_track(73, 1.84, 1),           # < 128 frames
            synth.synth_track(74, 1.9, 1),            # exactly enough for a few images
            synth.synth_track(75, 12.3, 3),
            np.zeros((44100 * 3, 1), np.int16)]       # digital silence: images exist, none survives
    got, st = ctx.index_pcm_batch(pcms)
    for pcm, g, s in zip(pcms, got, st):
        rc, sigs = oracle.fingerprint_pcm(pcm)
        assert s == rc
        assert bits_equal(g, sigs)
    assert st[0] == -5 and st[2] == -5 and len(got[5]) == 0 and st[5] == 0


def test_silent_gaps_are_compacted_in_order(ctx, oracle):
    pcm = synth.synth_track(81, 9.0, 1)
    pcm[2 * 44100:5 * 44100] = 0
    rc, sigs = oracle.fingerprint_pcm(pcm)
    got, st = ctx.index_pcm_batch([pcm])
    n_img = 1 + (1 + (len(pcm) // 8 - 2048) // 64 - 128) // 8
    assert 0 < len(sigs) < n_img                       # something was dropped
    assert bits_equal(got[0], sigs)


def test_clamped_rms_and_clipping(ctx, oracle):
    rng = np.random.RandomState(3)
    quiet = (rng.standard_normal((44100 * 3, 1)) * 2).astype(np.int16)
    loud = (np.sign(rng.standard_normal((44100 * 3, 2))) * 32767).astype(np.int16)
    for pcm in (quiet, loud):
        b = run_with_taps(ctx, pcm)
        s55 = oracle.resample(oracle.downmix(pcm))
        norm, rms = oracle.normalize(s55)
        assert np.float32(b.read(1, 0)[0]) == np.float32(rms)
        rc, sigs = oracle.fingerprint_pcm(pcm)
        got, st = b.signatures()
        assert bits_equal(got[0], sigs)
        b.close()


def test_randomized_signals_channels_and_lengths(ctx, oracle):
    # white noise at several levels, DC offsets, full-scale square waves, int16 extremes, 1..6 channels, ragged
    # lengths (not multiples of 8 / 64 / 4096): one ragged batch, every track compared with the oracle
    rng = np.random.RandomState(2024)
    pcms = []
    for k in range(14):
        ch = int(rng.randint(1, 7))
        n = int(rng.randint(81431, 3 * 81431)) + int(rng.randint(0, 9))
        kind = k % 5
        if kind == 0:
            x = rng.standard_normal((n, ch)) * rng.choice([30, 800, 9000, 30000])
        elif kind == 1:
            x = rng.standard_normal((n, ch)) * 500 + rng.choice([-20000, 15000])
        elif kind == 2:
            x = np.sign(np.sin(np.arange(n)[:, None] * rng.uniform(0.01, 0.3))) * 32767 * np.ones((1, ch))
        elif kind == 3:
            x = rng.choice([-32768, 32767, 0, 1, -1], size=(n, ch))
        else:
            x = synth.synth_track(900 + k, n / 44100.0 + 0.01, 1)[:n].astype(np.float64) * np.ones((1, ch)) * rng.uniform(0.2, 1.6)
        pcms.append(np.ascontiguousarray(np.clip(np.round(x), -32768, 32767).astype(np.int16)))
    got, st = ctx.index_pcm_batch(pcms)
    for pcm, g, s in zip(pcms, got, st):
        rc, sigs = oracle.fingerprint_pcm(pcm)
        assert s == rc
        assert bits_equal(g, sigs)


def tie_track(seed):
    """PCM whose 5512 Hz samples repeat every 2048 samples (= 32 frames): the 128-frame image consists of 4
    bit-identical repeats, so Haar coefficients come in groups of exactly equal magnitude and the 200th
    largest falls inside such a group (asserted on the oracle's coefficients in the test)."""
    rng = np.random.RandomState(seed)
    P = 16384
    t = np.arange(P)
    env = np.where((t // 2048) % 4 == 0, 1.0, 0.03)
    base = rng.standard_normal(P) * 6000 * env + 9000 * np.sin(2 * np.pi * t * (40 + (seed - 70) * 7) / P) * env[::-1]
    return np.tile(np.clip(np.round(base), -32768, 32767).astype(np.int16), 12)[:, None]


@pytest.mark.parametrize("seed", [70, 71])
def test_ties_at_rank_200_go_to_the_lower_index(ctx, oracle, seed):
    # rawfingerprints.c:90: qsort by |c|; glibc's merge sort is stable, so among equal magnitudes the lower
    # index is selected.  Raw fingerprints (which coefficients got a bit) and signatures must be identical.
    pcm = tie_track(seed)
    b = run_with_taps(ctx, pcm)
    norm, rms = oracle.normalize(oracle.resample(oracle.downmix(pcm)))
    rc, sigs, taps = oracle.fingerprint_samples(norm, taps=True)
    mags = np.sort(np.abs(taps["haar"]), axis=1)[:, ::-1]
    assert (mags[:, 199] == mags[:, 200]).sum() >= 10                      # the fixture does straddle the cut
    assert bits_equal(b.read(7, 0).reshape(-1, 4096), taps["haar"])
    assert bits_equal(b.read(5, 0).reshape(-1, 1024), taps["raw"])
    got, st = b.signatures()
    assert st == [0] and len(sigs) > 0 and bits_equal(got[0], sigs)
    b.close()


@pytest.mark.parametrize("direction", ["crescendo", "decrescendo", "steps"])
def test_window_maximum_changes_every_image(ctx, oracle, direction):
    # Group records (k15_groups.cu) are shared by the images of a frame group that have the same window maximum.
    # A steadily rising (falling) level gives every image its own maximum: all 16 slots of a group are in use;
    # level steps give runs of equal maxima in between.  Images, Haar coefficients and signatures must not notice.
    n = 44100 * 14
    t = np.arange(n) / 44100.0
    rng = np.random.RandomState(5)
    if direction == "crescendo":
        env = 0.02 + 0.9 * (t / t[-1]) ** 2
    elif direction == "decrescendo":
        env = 0.02 + 0.9 * (1 - t / t[-1]) ** 2
    else:
        env = 0.05 + 0.25 * np.floor(t / 1.3) % 4
    x = env * (np.sin(2 * np.pi * (500 + 300 * np.sin(2 * np.pi * 0.7 * t)) * t) + 0.2 * rng.standard_normal(n))
    pcm = np.clip(np.round(x * 20000), -32768, 32767).astype(np.int16)[:, None]
    b = run_with_taps(ctx, pcm)
    norm, rms = oracle.normalize(oracle.resample(oracle.downmix(pcm)))
    rc, sigs, taps = oracle.fingerprint_samples(norm, taps=True)
    assert rc == 0 and len(taps["images"]) > 100
    assert bits_equal(b.read(6, 0).reshape(-1, 4096), taps["images"])
    assert bits_equal(b.read(7, 0).reshape(-1, 4096), taps["haar"])
    assert bits_equal(b.read(5, 0).reshape(-1, 1024), taps["raw"])
    got, st = b.signatures()
    assert st == [0] and bits_equal(got[0], sigs)
    b.close()
    got2, st2 = ctx.index_pcm_batch([pcm])          # the production instantiation (no taps)
    assert bits_equal(got2[0], sigs)


def test_live_path_reuses_its_graph_and_follows_length_changes(ctx, oracle):
    """mnemo_index_samples keeps the batch of the last length with one captured CUDA graph (the ears loop calls it once a
    second with a window of constant length): repeated calls, new content, and a change of length all equal the oracle."""
    rng = np.random.RandomState(12)
    for n in (27562, 27562, 27562, 55125, 10176, 27562):
        s = np.clip(rng.standard_normal(n) * 0.3, -1, 1).astype(np.float32)
        rc, got = ctx.index_samples(s)
        rc_o, sigs = oracle.fingerprint_samples(s)
        assert rc == rc_o == 0 and bits_equal(got, sigs)
        assert ctx.live_graph
    assert ctx.index_samples(np.zeros(10175, np.float32))[0] == -5      # FILE_TOO_SMALL leaves the cached batch alone
    rc, got = ctx.index_samples(s)
    assert rc == 0 and bits_equal(got, sigs)


def test_from_samples_entry_point(ctx, oracle):
    # generate_fingerprint_from_samples (fingerprinting.c:81-109): no resample, no normalisation
    rng = np.random.RandomState(9)
    s = np.clip(rng.standard_normal(31444) * 0.3, -1, 1).astype(np.float32)
    rc, got = ctx.index_samples(s)
    rc_o, sigs = oracle.fingerprint_samples(s)
    assert rc == rc_o == 0 and bits_equal(got, sigs) and len(sigs) <= 42
    assert ctx.index_samples(s[:2047])[0] == -5
    assert ctx.index_samples(s[:10175])[0] == -5


def test_device_logf_equals_host_libm_on_its_whole_domain(ctx):
    # every float in [1, 256]: arguments of logf(1 + scaled), scaled in [0, 255]
    assert ctx.selftest_logf(0x3F800000, 0x43800000) == 0


def test_reciprocal_divisions_equal_ieee_division(ctx):
    # x / M_SQRT2 (haar.c:35-36) in double by reciprocal and x / 32767.0 (wav.c:373) in FP32: every one of
    # the 2^32 float dividends
    assert ctx.selftest_div(0, 0, 1 << 32) == 0
    assert ctx.selftest_div(1, 0, 1 << 32) == 0
    # the FP32 forms of x / M_SQRT2 and x / logf(256.0f) (spectralimages.c:57): every float of their domains
    # (0 or |x| >= 2^-103 resp. 2^-106; k2_fingerprint.cu shows the kernels stay inside)
    assert ctx.selftest_div(3, 0, 1 << 32) == 0
    assert ctx.selftest_div(4, 0, 1 << 32) == 0
    # (255*v) / max (spectralimages.c:53): 2^28 pseudo-random (v, max) pairs, double results bit-equal
    assert ctx.selftest_div(2, 12345, 1 << 28) == 0


def test_rerun_is_deterministic(ctx):
    pcm = synth.synth_track(91, 20.0, 2)
    b = ctx.batch([pcm])
    b.upload()
    b.run()
    first, _ = b.signatures()
    b.upload()
    b.run()
    second, _ = b.signatures()
    assert bits_equal(first[0], second[0])
    b.close()


def test_full_size_two_hour_stereo_properties(ctx, oracle):
    """BASELINE config #2 (2 h stereo, 1.27 GB PCM).  The oracle cannot run the whole path at this
    size in seconds, so: (a) the cheap stages are compared in full (this is where the sequential
    float32 sum of squares deviates 9 % from the true sum), (b) windows of images are compared
    with the oracle, (c) the signal is 10-minute periodic with a period that is a multiple of the
    image hop, so signatures must repeat with that period."""
    base = synth.synth_track(101, 600.0, 2)
    period = (len(base) // 4096) * 4096                  # whole number of image hops
    base = base[:period]
    pcm = np.ascontiguousarray(np.tile(base, (12, 1)))
    b = ctx.batch([pcm])
    b.upload()
    b.run()
    b.sync()
    s55 = oracle.resample(oracle.downmix(pcm))
    assert bits_equal(b.read(0, 0), s55)
    norm, rms = oracle.normalize(s55)
    assert np.float32(b.read(1, 0)[0]) == np.float32(rms)
    true = float(np.sum(s55.astype(np.float64) ** 2))
    assert abs(float(b.read(8, 0)[0]) - true) / true > 0.01      # the sequential sum is visibly not the true sum
    dense = b.read(4, 0).reshape(-1, 100)
    keep = b.read(3, 0)
    n_img = len(keep)
    assert n_img == 1 + (1 + (len(s55) - 2048) // 64 - 128) // 8
    for j0 in (0, n_img // 3, n_img - 40):
        w = norm[512 * j0: 512 * (j0 + 39) + 10176]
        rc, sigs, taps = oracle.fingerprint_samples(w, taps=True)
        assert rc == 0
        kept = keep[j0:j0 + 40].astype(bool)
        assert bits_equal(kept, taps["silence"] == 0)
        assert bits_equal(dense[j0:j0 + 40][kept], sigs)
    k = period // 4096
    assert bits_equal(keep[:n_img - k], keep[k:])
    both = keep[:n_img - k].astype(bool)
    assert bits_equal(dense[:n_img - k][both], dense[k:][both])
    got, st = b.signatures()
    assert st == [0] and bits_equal(got[0], dense[keep.astype(bool)])
    b.close()


def test_staged_api_reuse_and_error_paths(ctx, oracle):
    from mnemophonix_b200 import capi
    # empty batch
    got, st = ctx.index_pcm_batch([])
    assert got == [] and st == []
    # zero-length track next to a real one
    a = synth.synth_track(201, 3.0, 1)
    got, st = ctx.index_pcm_batch([np.zeros((0, 2), np.int16), a])
    assert st[0] == -5 and len(got[0]) == 0 and bits_equal(got[1], oracle.fingerprint_pcm(a)[1])
    # a batch object is reusable: upload different audio of the same shape, run again
    b = ctx.batch([a])
    b.upload()
    b.run()
    first, _ = b.signatures()
    a2 = synth.synth_track(202, 3.0, 1)
    b.upload([a2])
    b.run()
    second, _ = b.signatures()
    assert bits_equal(first[0], oracle.fingerprint_pcm(a)[1]) and bits_equal(second[0], oracle.fingerprint_pcm(a2)[1])
    # shape mismatch on upload and too small an output buffer are reported, not ignored
    with pytest.raises(capi.MnemoError):
        b.upload([a2[:-8]])
    out = np.empty((1, 100), np.uint8)
    n = np.zeros(1, np.uint32)
    stt = np.zeros(1, np.int32)
    rc = capi.lib().mnemo_batch_download(b.h, out.ctypes.data, 1, n.ctypes.data, stt.ctypes.data)
    assert rc == -1 and int(n[0]) == len(second[0])
    # taps are off by default: asking for them is an error, not garbage
    with pytest.raises(capi.MnemoError):
        b.read(6, 0)
    b.close()


def test_corpus_streaming_workload_is_batch_independent_and_equals_the_oracle(oracle):
    """workloads.stream_index (BASELINE configs[2] / [4] on the bench line) on a small corpus: 14 tracks cycled from a ring
    of 3, ragged batches of 4 in flight on two streams, two passes; every track's signatures equal the ones its ring track
    produced alone (count + CRC-32), and that single-track run equals the oracle."""
    import torch
    from mnemophonix_b200 import workloads

    def checker(pcm, sigs):
        rc, want = oracle.fingerprint_pcm(pcm)
        return rc == 0 and np.array_equal(want, sigs)

    out = workloads.stream_index(0, torch.device("cuda", 0), 0, 1, None, n_tracks=14, seconds=12.0, per_batch=4, distinct=3,
                                 passes=2, checker=checker)
    assert out["tracks_crc_checked"] == 28 and out["tracks_differing_from_single_track_run"] == 0
    assert out["ring_track_equals_oracle"] is True
    assert out["signatures"] > 0 and out["audio_hours_per_s"] > 0
