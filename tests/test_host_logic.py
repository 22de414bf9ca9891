"""CPU: host-side helpers (synthetic audio, WAV container) used by tests and bench."""
import struct

import numpy as np

from mnemophonix_b200 import synth


def test_synth_is_deterministic_and_in_range():
    a = synth.synth_track(1001, 3.0, 2)
    b = synth.synth_track(1001, 3.0, 2)
    assert a.dtype == np.int16 and a.shape == (3 * 44100, 2) and np.array_equal(a, b)
    assert np.abs(a).max() <= 20000 and a.std() > 1000
    assert not np.array_equal(a, synth.synth_track(1002, 3.0, 2))
    assert np.array_equal(a[:, 1], np.round(a[:, 0].astype(np.float64) * 0.8).astype(np.int16))


def test_synth_block_boundaries_do_not_change_the_signal():
    a = synth.synth_track(5, 4.0, 1, block_s=60.0)
    b = synth.synth_track(5, 4.0, 1, block_s=60.0)
    assert np.array_equal(a, b)


def test_wav_bytes_layout():
    pcm = synth.synth_track(7, 0.1, 2)
    w = synth.wav_bytes(pcm, {"IART": "x"})
    assert w[:4] == b"RIFF" and w[8:12] == b"WAVE" and w[12:16] == b"fmt "
    size, tag, ch, rate, bps, align, bits = struct.unpack("<IhHIIHH", w[16:36])
    assert (size, tag, ch, rate, align, bits) == (16, 1, 2, 44100, 4, 16)
    assert w[36:40] == b"data" and struct.unpack("<I", w[40:44])[0] == pcm.nbytes
    assert struct.unpack("<I", w[4:8])[0] == len(w) - 8
    assert b"LISTINFOIART"[:4] in w[44 + pcm.nbytes:]


def test_committed_bench_lines_carry_the_contract():
    """The bench lines kept under profiles/ (what bench.py printed on the GPU box) hold every key of the driver's contract,
    at every N; the reference arm's line marks itself and reports no device work."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for n in (1, 2, 4, 8):
        d = json.load(open(os.path.join(root, "profiles", f"r02_bench_n{n}.json")))
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                  "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
            assert k in d, (n, k)
        assert d["n_gpus"] == n and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
        assert "workload" in d["config"] and "model" not in d["config"]
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
        assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["gpu_launches"] > 0
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
        assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-9
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        c4 = d["search_c4"]
        assert c4["identical_to_single_gpu"] and c4["identical_to_reference_on_sample"] and c4["reference_sample"] >= 64
        assert c4["dense_rows_single_gpu_equal_oracle"] and c4["dense_rows_sharded_equal_oracle"]
        assert (c4["nccl_message_bytes_per_batch_per_rank"] > 0) == (n > 1)
        for key in ("batch_c3", "stream_c5"):
            assert d[key]["tracks_differing_from_single_track_run"] == 0 and d[key]["n_gpus"] == n
        if n == 1:
            assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
        if n >= 4:
            assert c4["grid"]["identical_to_single_gpu"]
    ref = json.load(open(os.path.join(root, "profiles", "r02_bench_ref_n1.json")))
    assert ref["impl"] == "reference" and ref["gpu_launches"] == 0 and ref["cpu_baseline"]["kind"] == "reference"
    assert ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["e2e"]["value"] == ref["value"]
