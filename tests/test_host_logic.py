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
