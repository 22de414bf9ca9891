"""CPU: pin the oracle (oracle/mnemo_oracle.c) against golden vectors produced by the unmodified
reference (tools/gen_golden.py) and against the reference's own README known answer."""
import numpy as np


def bits_equal(a, b):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    return a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8))


def test_tables(oracle, golden):
    assert bits_equal(oracle.hann(), golden["hann"])
    e = oracle.band_edges()
    # SURVEY.md App. B (computed with glibc from logbins.c:44-55)
    assert e.tolist() == [118, 125, 133, 140, 149, 157, 167, 177, 187, 198, 210, 222, 235, 249, 264, 280, 296,
                          314, 332, 352, 373, 395, 418, 443, 469, 497, 526, 558, 591, 625, 662, 702, 743]
    h = oracle.taps()
    assert h[15] == np.float32(0.125) and abs(float(h.sum()) - 0.9666) < 1e-3
    perm = oracle.permutations()
    assert perm.shape == (100, 255) and perm.max() < 8192
    assert all(len(set(row.tolist())) == 255 for row in perm)       # every row: 255 distinct bit positions
    assert len(set(perm.reshape(-1).tolist())) == 7823


def test_twiddle_table_reproduces_every_stage(oracle):
    wr, wi = oracle.twiddles()
    for l in (2, 4, 64, 512, 2048):
        for k in range(0, l // 2, max(1, l // 32)):
            ang = np.float32(-2.0 * k * np.pi / l)
            idx = k * (2048 // l)
            assert np.float32(-2.0 * idx * np.pi / 2048) == ang    # identical angle => identical cosf/sinf


def test_readme_known_answer(oracle, golden):
    # reference README.md:52-56: 31444 samples -> 42 spectral images
    assert oracle.lib.mo_n_images(oracle.lib.mo_n_frames(31444)) == 42 == int(golden["readme_n_images"][0])


def test_stages_against_golden(oracle, golden):
    for name in ("a", "b"):
        pcm = golden[f"{name}_pcm"]
        s44 = oracle.downmix(pcm)
        s55 = oracle.resample(s44)
        assert bits_equal(s55, golden[f"{name}_s5512"])
        norm, rms = oracle.normalize(s55)
        assert bits_equal(norm[:4096], golden[f"{name}_norm_head"])
        bins = np.stack([oracle.frame_bins(norm[64 * f:64 * f + 2048]) for f in range(len(golden[f"{name}_bins_head"]))])
        assert bits_equal(bins, golden[f"{name}_bins_head"])
        rc, sigs, taps = oracle.fingerprint_samples(norm, taps=True)
        assert rc == 0
        pick = golden[f"{name}_pick"]
        assert bits_equal(taps["images"][pick], golden[f"{name}_images"])
        assert bits_equal(taps["haar"][pick], golden[f"{name}_haar"])
        assert bits_equal(taps["raw"], golden[f"{name}_raw"])
        assert bits_equal(taps["silence"], golden[f"{name}_silence"])
        assert bits_equal(sigs, golden[f"{name}_sigs"])
        rc, sigs2 = oracle.fingerprint_pcm(pcm)
        assert rc == 0 and bits_equal(sigs2, golden[f"{name}_sigs"])


def test_logf_restatement_matches_host_libm_exhaustively(oracle):
    # every float in [1, 256]: the whole domain of spectralimages.c:57's logf(1 + scaled)
    assert oracle.logf_check(0x3F800000, 0x43800000, 1) == 0
    assert oracle.logf_check(0x3F800000, 0x43800000, 0) == 0


def test_too_small_inputs(oracle):
    rc, sigs = oracle.fingerprint_samples(np.zeros(2047, np.float32))
    assert rc == -5 and len(sigs) == 0
    rc, sigs = oracle.fingerprint_samples(np.zeros(10175, np.float32))     # 127 frames
    assert rc == -5
    rc, sigs = oracle.fingerprint_samples(np.zeros(10176, np.float32))     # 128 frames, but silent
    assert rc == 0 and len(sigs) == 0


def test_search_against_golden(oracle):
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "search_golden.npz"))
    n = int(g["n_tracks"][0])
    tracks = [g[f"track{i}"] for i in range(n)]
    db = oracle.make_db(tracks)
    res = []
    for i in range(n + 1):
        best, sc, L, D = oracle.search(db, g[f"query{i}"])
        res.append(best)
    best, sc, L, D = oracle.search(db, tracks[1])
    res.append(best)
    assert res == g["results"].tolist()
    assert oracle.get_matches(db, g["gm0_sig"]) == [tuple(x) for x in g["gm0"].tolist()]
    assert oracle.get_matches(db, g["gm1_sig"]) == [tuple(x) for x in g["gm1"].tolist()]


def test_comparator_is_not_a_strict_weak_order(oracle):
    # A=(avg 40,n 5), B=(37,10), C=(34,15): B beats A, C beats B, yet A beats C.
    score = [40.0 * 5, 37.0 * 10, 34.0 * 15]
    n = [5, 10, 15]
    # whatever order qsort leaves them in, selection must pick the greatest qualifying average = entry 0
    assert oracle.select(score, n) == 0
    assert oracle.select([0.0, 0.0], [0, 0]) == -7
    assert oracle.select([29.0 * 12], [12]) == -7          # avg below 30
    assert oracle.select([36.0 * 5], [5]) == 0             # avg >= 35 with 5 matches qualifies
    assert oracle.select([34.0 * 9], [9]) == -7            # 9 matches, avg < 35
