"""Dev tool: wall time of `mnemophonix index <2 h WAV>` (our CLI, whole process incl. CUDA context creation) and of
the reference library's generate_fingerprint on the same file."""
import os, subprocess, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mnemophonix_b200 import synth
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 7200.0
base = synth.synth_track(1, min(secs, 600.0), 2)
pcm = np.ascontiguousarray(np.tile(base, (int(round(secs / min(secs, 600.0))), 1)))
path = "/dev/shm/mnemo_cli_2h.wav"
synth.write_wav(path, pcm)
cli = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mnemophonix_b200", "mnemophonix")
for it in range(3):
    t = time.perf_counter()
    r = subprocess.run([cli, "index", path], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=dict(os.environ, MNEMO_TRACE="1"))
    if it == 0: print(r.stderr.decode()[-900:])
    dt = time.perf_counter() - t
    print("ours: rc %d, %.3f s wall, %d bytes of text, %.1f audio-h/s" % (r.returncode, dt, len(r.stdout), secs / 3600 / dt))
# several files in one process: the context is created once
paths = [path] + ["/dev/shm/mnemo_cli_2h_%d.wav" % i for i in range(1, 4)]
for q in paths[1:]:
    os.link(path, q) if not os.path.exists(q) else None
t = time.perf_counter()
r = subprocess.run([cli, "index"] + paths, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=dict(os.environ, MNEMO_TRACE="1"))
dt = time.perf_counter() - t
print("ours, 4 files in one process: rc %d, %.3f s wall, %.1f audio-h/s" % (r.returncode, dt, 4 * secs / 3600 / dt))
print(r.stderr.decode()[-700:])
for q in paths[1:]:
    os.remove(q)
if "--ref" in sys.argv:
    from oracle import ref as refmod
    if refmod.available("O2"):
        rr = refmod.Ref("O2")
        t = time.perf_counter(); rc, sigs, *_ = rr.fingerprint_wav(path); dt = time.perf_counter() - t
        print("reference library (-O2, one process, 8 threads): %.2f s, %d signatures" % (dt, len(sigs)))
os.remove(path)
