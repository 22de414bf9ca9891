"""Profiling driver: BASELINE.json configs[1] (one 2 h stereo track, 1.27 GB of PCM) indexed `--runs` times, device-resident,
in the exact mode or (--fast) the tolerance mode of K1; nothing else runs, so that an ncu capture of this command holds
exactly the index kernels of that workload.

    python tools/profile_index.py [--fast] [--runs 3] [--seconds 7200]
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mnemophonix_b200 import capi
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--fast", action="store_true")
ap.add_argument("--runs", type=int, default=3)
ap.add_argument("--seconds", type=float, default=7200.0)
args = ap.parse_args()
stream = torch.cuda.Stream()
ctx = capi.Context(0, stream.cuda_stream, fast_spectra=args.fast)
pcm = bench.make_track(args.seconds, 2, 1)
pin = torch.from_numpy(pcm).pin_memory()
b = ctx.batch([(pin.data_ptr(), pcm.shape[0], 2)])
b.upload(); b.sync()
b.set_timing(True)
for _ in range(args.runs):
    b.run()
b.sync()
print(b.get_timing(), "launches per run", b.launches(), "fast" if args.fast else "exact")
b.close(); ctx.close()
