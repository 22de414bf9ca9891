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


def synth_tracks_torch(track_ids, seconds: float, dev):
    """The recipe of synth_track evaluated with torch on `dev` for a batch of tracks: int16 [B, n] (mono).  Track t is
    a function of t alone (its own generator seeds) but not sample-identical to synth_track(t): different random
    streams.  Used where tens of thousands of tracks are needed in bench time (BASELINE config #4)."""
    import math
    import torch
    n = int(round(seconds * RATE))
    B = len(track_ids)
    idx = torch.arange(n, device=dev, dtype=torch.float64)
    t = idx / RATE
    K = int(math.ceil(seconds / 0.08)) + 1
    par = torch.empty((B, 3), dtype=torch.float64)
    seg = torch.empty((B, 3, K, 3), dtype=torch.float64)     # duration, frequency, gain
    noise = torch.empty((B, n), device=dev, dtype=torch.float32)
    for i, tid in enumerate(track_ids):
        g = torch.Generator().manual_seed(1000 + int(tid))
        u = torch.rand(3, generator=g, dtype=torch.float64)
        par[i] = torch.stack([2.0 + 3.0 * u[0], 250.0 + 150.0 * u[1], 1500.0 + 700.0 * u[2]])
        v = torch.rand((3, K, 3), generator=g, dtype=torch.float64)
        seg[i, :, :, 0] = 0.08 + 0.52 * v[:, :, 0]
        seg[i, :, :, 1] = 300.0 + 1700.0 * v[:, :, 1]
        seg[i, :, :, 2] = 0.3 * v[:, :, 2]
        gg = torch.Generator(device=dev).manual_seed(77000 + int(tid))
        noise[i] = torch.randn(n, generator=gg, device=dev, dtype=torch.float32)
    par = par.to(dev); seg = seg.to(dev)
    period, f0, f1 = par[:, 0:1], par[:, 1:2], par[:, 2:3]
    tau = torch.remainder(t.unsqueeze(0), period)
    x = 0.35 * torch.sin(2.0 * math.pi * (f0 * tau + 0.5 * (f1 - f0) / period * tau * tau))
    del tau
    ramp = int(0.010 * RATE)
    for vce in range(3):
        dur = seg[:, vce, :, 0]
        start = torch.cumsum(dur, dim=1) - dur                                    # [B, K] seconds
        s0 = torch.floor(start * RATE)                                            # start sample
        ln = torch.floor(dur * RATE)
        which = torch.searchsorted(s0.contiguous(), idx.unsqueeze(0).expand(B, n).contiguous(), right=True) - 1
        which.clamp_(0, K - 1)
        k = idx.unsqueeze(0) - torch.gather(s0, 1, which)
        lnw = torch.gather(ln, 1, which)
        inside = k < lnw
        env = torch.clamp(torch.minimum(k, lnw - 1 - k) / ramp, max=1.0)
        fr = torch.gather(seg[:, vce, :, 1], 1, which)
        gn = torch.gather(seg[:, vce, :, 2], 1, which)
        x += torch.where(inside, gn * env * torch.sin(2.0 * math.pi * fr * k / RATE), torch.zeros((), device=dev, dtype=torch.float64))
        del which, k, lnw, inside, env, fr, gn
    x += 0.05 * noise.double()
    x.clamp_(-1.0, 1.0)
    return torch.round(x * 20000.0).to(torch.int16)


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
