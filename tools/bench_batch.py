"""BASELINE config #3 / #5 style batch indexing on one GPU: many 4-minute mono tracks, host buffers in,
signatures out, two batches in flight so uploads overlap kernels (the streaming pattern of config #5).

    python tools/bench_batch.py [--tracks 130] [--seconds 240] [--per-batch 65] [--rounds 4]"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mnemophonix_b200 import capi, synth

ap = argparse.ArgumentParser()
ap.add_argument("--tracks", type=int, default=130)
ap.add_argument("--seconds", type=float, default=240.0)
ap.add_argument("--per-batch", type=int, default=65)
ap.add_argument("--rounds", type=int, default=4)      # the corpus is streamed this many times (config #5 flavour)
ap.add_argument("--distinct", type=int, default=13)   # distinct synthetic tracks (the rest are copies: host RAM/time)
args = ap.parse_args()

base = [synth.synth_track(1000 + i, args.seconds, 1) for i in range(args.distinct)]
n_frames = base[0].shape[0]
pinned = [torch.from_numpy(b).pin_memory() for b in base]
tracks = [(pinned[i % args.distinct].data_ptr(), n_frames, 1) for i in range(args.tracks)]
batches = [tracks[i:i + args.per_batch] for i in range(0, args.tracks, args.per_batch)]
lanes = []
for k in range(2):
    st = torch.cuda.Stream()
    ctx = capi.Context(0, st.cuda_stream)
    lanes.append((ctx, {}))

def get_batch(lane, key, tr):
    ctx, cache = lanes[lane]
    if key not in cache:
        b = ctx.batch(tr)
        cache[key] = (b, torch.empty((max(b.cap, 1), 100), dtype=torch.uint8).pin_memory())
    return cache[key]

def run(rounds):
    jobs = [(r, i) for r in range(rounds) for i in range(len(batches))]
    pending = None
    total = 0
    for j, (r, i) in enumerate(jobs):
        b, out = get_batch(j % 2, len(batches[i]), batches[i])
        b.upload(batches[i]); b.run()
        if pending is not None:
            _, n, _ = pending[0].download(out_ptr=pending[1].data_ptr()); total += int(n.sum())
        pending = (b, out)
    _, n, _ = pending[0].download(out_ptr=pending[1].data_ptr()); total += int(n.sum())
    return total

run(1)
torch.cuda.synchronize()
t0 = time.perf_counter()
sigs = run(args.rounds)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
hours = args.rounds * args.tracks * args.seconds / 3600.0
pcm_bytes = args.rounds * args.tracks * n_frames * 2
# device-resident rate of one batch for comparison
b, out = get_batch(0, len(batches[0]), batches[0])
b.upload(batches[0]); b.sync()
for _ in range(2): b.run()
b.sync()
t1 = time.perf_counter()
for _ in range(5): b.run()
b.sync()
dev = (time.perf_counter() - t1) / 5
print(json.dumps({"config": f"{args.tracks} x {args.seconds:.0f} s mono tracks per round, {args.rounds} rounds, batches of {args.per_batch}, "
                            "2 batches in flight (overlapped H2D)", "audio_hours": hours, "e2e_audio_hours_per_s": hours / dt,
                  "e2e_s": dt, "h2d_GBps": pcm_bytes / dt / 1e9, "signatures": sigs,
                  "device_resident_audio_hours_per_s": len(batches[0]) * args.seconds / 3600.0 / dev,
                  "device_ms_per_batch": dev * 1e3}))
