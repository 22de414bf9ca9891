"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, tables pre-warmed).

The reference ships no fixtures, so the golden vectors are outputs of the
reference library itself, run in this container:

    make -C oracle ref && python tools/gen_golden.py

index_golden.npz — a 4 s stereo and a 6 s mono clip (PCM stored), with the reference's 5512 Hz
    samples, normalised samples, first frames' band energies, spectral images / Haar / raw bits of a
    few images (full tensors would be MBs) and ALL signatures;
search_golden.npz — a 5-entry database (signatures from the reference), 5 query clips, the
    reference's search() result per query and get_matches() lists for two signatures.
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mnemophonix_b200 import synth  # noqa: E402
from oracle.ref import Ref  # noqa: E402

r = Ref("O2")
out = {}
for name, seed, secs, ch in (("a", 11, 4.0, 2), ("b", 12, 6.0, 1)):
    pcm = synth.synth_track(seed, secs, ch)
    s44 = r.downmix(pcm)
    s55 = r.resample(s44)
    norm = r.normalize(s55)
    st = r.stages_from_samples(norm)
    # end-to-end through a WAV file and generate_fingerprint as well
    p = tempfile.mktemp(suffix=".wav")
    synth.write_wav(p, pcm)
    rc, sigs_wav, _, _, _ = r.fingerprint_wav(p)
    os.remove(p)
    assert rc == 0 and np.array_equal(sigs_wav, st["sigs"])
    nf = 1 + (len(norm) - 2048) // 64
    bins = np.stack([r.bins(*r.fft(norm[64 * f:64 * f + 2048] * r.hann())) for f in range(min(nf, 160))])
    pick = [0, len(st["images"]) // 2, len(st["images"]) - 1]
    out[f"{name}_pcm"] = pcm
    out[f"{name}_s5512"] = s55
    out[f"{name}_norm_head"] = norm[:4096]
    out[f"{name}_bins_head"] = bins
    out[f"{name}_pick"] = np.array(pick)
    out[f"{name}_images"] = st["images"][pick]
    out[f"{name}_haar"] = st["haar"][pick]
    out[f"{name}_raw"] = st["raw"]
    out[f"{name}_silence"] = st["silence"]
    out[f"{name}_sigs"] = st["sigs"]
out["hann"] = r.hann()
# README.md:52-56 known answer: 31444 samples -> 42 images
s = np.random.RandomState(0).uniform(-0.5, 0.5, 31444).astype(np.float32)
st = r.stages_from_samples(s)
out["readme_n_images"] = np.array([len(st["images"])])
os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "index_golden.npz"), **out)

# ---- search ----
tracks, queries = [], []
for tid in range(5):
    pcm = synth.synth_track(2000 + tid, 16.0, 1)
    rc, sg = r.fingerprint_pcm(pcm)
    tracks.append(sg)
    rc, q = r.fingerprint_pcm(pcm[4 * 44100 + 1234: 9 * 44100 + 1234])
    queries.append(q)
rng = np.random.RandomState(5)
queries.append(rng.randint(0, 60, size=(20, 100)).astype(np.uint8))   # unrelated query -> no match
p = tempfile.mktemp()
with open(p, "w") as f:
    for i, sg in enumerate(tracks):
        f.write(f"track{i}.wav\n\n\n\n{len(sg)}\n")
        for row in sg:
            f.write("".join(f"{b:02x}" for b in row) + "\n")
rc, pidx = r.load_index(p)
assert rc == 0
lsh = r.lib.create_hash_tables(pidx)
res = [r.search(q, pidx, lsh) for q in queries] + [r.search(tracks[1], pidx, lsh)]
gm0 = np.array(r.get_matches(lsh, queries[0][3]), np.uint32).reshape(-1, 2)
gm1 = np.array(r.get_matches(lsh, tracks[2][7]), np.uint32).reshape(-1, 2)
os.remove(p)
sout = {"n_tracks": np.array([5]), "results": np.array(res, np.int32), "gm0": gm0, "gm1": gm1,
        "gm0_sig": queries[0][3], "gm1_sig": tracks[2][7]}
for i, sg in enumerate(tracks):
    sout[f"track{i}"] = sg
for i, q in enumerate(queries):
    sout[f"query{i}"] = q
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "search_golden.npz"), **sout)
print("search results", res)
print({k: v.shape for k, v in out.items()})
