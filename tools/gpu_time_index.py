"""Dev tool: time the device-resident index path on a long stereo track (CUDA events on torch's stream)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mnemophonix_b200 import capi, synth

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 7200.0
ch = int(sys.argv[2]) if len(sys.argv) > 2 else 2
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
base = synth.synth_track(1, min(secs, 600.0), ch)
reps = int(round(secs / min(secs, 600.0)))
pcm = np.ascontiguousarray(np.tile(base, (reps, 1)))
print("pcm", pcm.shape, pcm.nbytes / 1e9, "GB")
torch.cuda.init()
st = torch.cuda.current_stream()
ctx = capi.Context(0, st.cuda_stream)
pin = torch.from_numpy(pcm).pin_memory()
b = ctx.batch([(pin.data_ptr(), pcm.shape[0], ch)])
t = time.time(); b.upload(); b.sync(); print("upload s", time.time() - t, "GB/s", pcm.nbytes / 1e9 / (time.time() - t))
for _ in range(3):
    b.run()
b.sync()
b.set_timing(True)
t = time.perf_counter()
for i in range(steps):
    b.run()
b.sync()
ms = (time.perf_counter() - t) * 1e3 / steps
stage, n = b.get_timing()
hours = secs / 3600.0
print("ms/run (host clock, back to back)", ms, "stage ms", stage, "sum", sum(stage.values()), "audio-h/s",
      hours / (sum(stage.values()) / 1e3), "launches", b.launches())
out, n, stt = b.download(); print("n_sigs", n, stt)
