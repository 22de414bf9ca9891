"""Deterministic synthetic 44.1 kHz PCM used by tests and bench.

Signal = repeating linear chirp + 3 voices of tone bursts + white noise, clipped and
quantised to int16.  The reference has no fixtures of its own, so every
parity input in this repo comes from here (seed = 1000 + track_id by convention).
"""
from __future__ import annotations

import struct

import numpy as np

RATE = 44100


def synth_track(seed: int, seconds: float, channels: int = 1, block_s: float = 60.0) -> np.ndarray:
    """Return int16 PCM of shape [n_frames, channels] (C-contiguous, interleaved)."""
    n = int(round(seconds * RATE))
    rng = np.random.RandomState(seed)
    period = rng.uniform(2.0, 5.0)
    f0 = rng.uniform(250.0, 400.0)
    f1 = rng.uniform(1500.0, 2200.0)

    # tone-burst schedule for 3 voices: (start_sample, n_samples, freq, gain)
    bursts = []
    for _voice in range(3):
        t = 0.0
        while t < seconds:
            dur = rng.uniform(0.08, 0.6)
            freq = rng.uniform(300.0, 2000.0)
            gain = rng.uniform(0.0, 0.3)
            bursts.append((int(t * RATE), int(dur * RATE), freq, gain))
            t += dur
    bursts.sort()

    out = np.empty(n, dtype=np.int16)
    blk = int(block_s * RATE)
    ramp = int(0.010 * RATE)
    bi = 0
    for start in range(0, n, blk):
        stop = min(n, start + blk)
        idx = np.arange(start, stop, dtype=np.float64)
        t = idx / RATE
        # (i) repeating linear chirp
        tau = np.mod(t, period)
        phase = 2.0 * np.pi * (f0 * tau + 0.5 * (f1 - f0) / period * tau * tau)
        x = 0.35 * np.sin(phase)
        # (ii) tone bursts overlapping this block
        while bi < len(bursts) and bursts[bi][0] + bursts[bi][1] <= start:
            bi += 1
        j = bi
        while j < len(bursts) and bursts[j][0] < stop:
            s0, ln, freq, gain = bursts[j]
            a = max(s0, start)
            b = min(s0 + ln, stop)
            if b > a and ln > 2 * ramp:
                k = np.arange(a - s0, b - s0, dtype=np.float64)
                env = np.minimum(1.0, np.minimum(k, ln - 1 - k) / ramp)
                x[a - start:b - start] += gain * env * np.sin(2.0 * np.pi * freq * k / RATE)
            j += 1
        # (iii) noise
        x += 0.05 * rng.standard_normal(stop - start)
        np.clip(x, -1.0, 1.0, out=x)
        out[start:stop] = np.round(x * 20000.0).astype(np.int16)

    if channels == 1:
        return out.reshape(n, 1)
    pcm = np.empty((n, channels), dtype=np.int16)
    pcm[:, 0] = out
    for c in range(1, channels):
        pcm[:, c] = np.round(out.astype(np.float64) * (0.8 ** c)).astype(np.int16)
    return pcm


def wav_bytes(pcm: np.ndarray, info: dict | None = None) -> bytes:
    """Canonical RIFF/WAVE (PCM tag 1, 44.1 kHz, 16-bit, 16-byte fmt chunk) around `pcm`.

    `info` may hold IART / INAM / IPRD strings, written as a trailing LIST/INFO chunk
    (the layout the reference's wav.c:188-255 parses).
    """
    pcm = np.ascontiguousarray(pcm, dtype="<i2")
    n, ch = pcm.shape
    data = pcm.tobytes()
    fmt = struct.pack("<4sIhHIIHH", b"fmt ", 16, 1, ch, RATE, RATE * ch * 2, ch * 2, 16)
    tail = b""
    if info:
        body = b"INFO"
        for key in ("IART", "INAM", "IPRD"):
            if key in info:
                s = info[key].encode()
                body += key.encode() + struct.pack("<I", len(s)) + s
        tail = b"LIST" + struct.pack("<I", len(body)) + body
    payload = b"WAVE" + fmt + b"data" + struct.pack("<I", len(data)) + data + tail
    return b"RIFF" + struct.pack("<I", len(payload)) + payload


def write_wav(path: str, pcm: np.ndarray, info: dict | None = None) -> None:
    with open(path, "wb") as f:
        f.write(wav_bytes(pcm, info))
