#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 index path (BASELINE.json config #2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--seconds S]

Workload: index ONE 2-hour synthetic 44.1 kHz 16-bit STEREO track (1.27 GB of PCM, 77 500 spectral
images) per GPU.  A "step" is one pass of the whole hot path (PCM -> signatures) over that track.
  value   audio-hours fingerprinted per second, whole job, PCM already resident in HBM
          (CUDA events on the launching stream, max over ranks);
  e2e     the same through the C ABI with HOST buffers: pinned-host -> device copy of the PCM,
          kernels, device -> host copy of the signatures, every step;
  roofline  the dominant kernel's algorithmic bytes / its measured duration against measured HBM;
  cpu_baseline  the unmodified reference (oracle/_ref, -O2 build) on this box's host cores on a
          bounded sample of the same workload (N = 1, rank 0 only).
N > 1 (torchrun, one rank per GPU): indexing shards whole tracks across GPUs with no data-path
collective; every rank indexes its own 2-hour track (weak scaling), value = N tracks / max time.
--impl reference: the reference's own CPU implementation timed on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "audio-hours fingerprinted/s/GPU at 1/2/4/8 B200 (% HBM roofline); search queries/s"
UNIT = "audio-hours/s"
BYTES_PER_AUDIO_SECOND = {1: 44100 * 2 * 1 + 100 * 5512.5 / 512, 2: 44100 * 2 * 2 + 100 * 5512.5 / 512}


def make_track(seconds: float, channels: int, seed: int) -> np.ndarray:
    """2 h of synthetic audio: a 10-minute synthetic take (SURVEY.md §8d generator) tiled."""
    from mnemophonix_b200 import synth
    take = min(seconds, 600.0)
    base = synth.synth_track(seed, take, channels)
    reps = int(np.ceil(seconds / take))
    pcm = np.tile(base, (reps, 1))[: int(round(seconds * 44100))]
    return np.ascontiguousarray(pcm)


def make_worst_content_track(seconds: float, channels: int, seed: int) -> np.ndarray:
    """Content that defeats the sharing of group records (k15_groups.cu): a steady tone whose level doubles every second
    for ten seconds and then drops back (a saw-tooth crescendo, 60 dB per tooth).  While the level rises, the loudest
    frame of every spectral image is its newest one, so each image has its own window maximum and an 8-frame group
    needs a record for nearly every one of the 16 images that contain it, instead of the ~5 a steady-level take needs
    (the slots in use are reported)."""
    n = int(round(min(seconds, 600.0) * 44100))          # 60 saw-teeth, then tiled
    t = np.arange(n, dtype=np.float64) / 44100.0
    rng = np.random.RandomState(seed)
    env = 0.0009 * np.exp2(t % 10.0)                     # 0.0009 .. 0.92
    x = env * (np.sin(2 * np.pi * 1000.0 * t) + 0.5 * np.sin(2 * np.pi * 640.0 * t) + 0.05 * rng.standard_normal(n))
    mono = np.clip(np.round(x * 20000), -32768, 32767).astype(np.int16)
    take = np.stack([mono] + [np.round(mono * 0.8).astype(np.int16)] * (channels - 1), axis=1) if channels > 1 else mono[:, None]
    reps = int(np.ceil(seconds * 44100 / n))
    return np.ascontiguousarray(np.tile(take, (reps, 1))[: int(round(seconds * 44100))])


def dropin_section(pcm: np.ndarray, seconds: float, device: int, reps: int = 3):
    """End to end through the DROP-IN library: generate_fingerprint(path) of libmnemophonix.so (the reference's own entry
    point, fingerprinting.h:26) on a WAV file in tmpfs, in a warm process: header parse, threaded pread of the data
    chunk into a ring of pinned buffers, host -> device copies, kernels, signatures back."""
    import ctypes as C
    from mnemophonix_b200 import synth
    tmpdir = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else tempfile.gettempdir()
    path = os.path.join(tmpdir, f"mnemo_dropin_{os.getpid()}.wav")
    synth.write_wav(path, pcm)
    os.environ["MNEMO_DEVICE"] = str(device)
    lib = C.CDLL(os.path.join(ROOT, "mnemophonix_b200", "libmnemophonix.so"))

    class Signatures(C.Structure):
        _fields_ = [("n_signatures", C.c_uint), ("signatures", C.c_void_p)]
    lib.generate_fingerprint.argtypes = [C.c_char_p, C.POINTER(C.POINTER(Signatures)), C.c_void_p, C.c_void_p, C.c_void_p]
    lib.free_signatures.argtypes = [C.POINTER(Signatures)]
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(2)
    times, n_sigs = [], 0
    try:
        os.dup2(devnull, 2)                      # the reference's progress lines go to stderr on every call
        for it in range(reps + 1):               # first call: CUDA context of the library + pinned ring
            p = C.POINTER(Signatures)()
            t0 = time.perf_counter()
            rc = lib.generate_fingerprint(path.encode(), C.byref(p), None, None, None)
            dt = time.perf_counter() - t0
            assert rc == 0, rc
            n_sigs = int(p.contents.n_signatures)
            lib.free_signatures(p)
            if it:
                times.append(dt)
    finally:
        os.dup2(saved, 2)
        os.close(saved)
        os.close(devnull)
        if os.path.exists(path):
            os.remove(path)
    dt = float(np.median(times))
    return {"value": (seconds / 3600.0) / dt, "unit": UNIT, "ms_per_call": dt * 1e3, "file_bytes": int(pcm.nbytes),
            "read_plus_upload_GBps_lower_bound": pcm.nbytes / dt / 1e9, "signatures": n_sigs,
            "note": "generate_fingerprint(path) through libmnemophonix.so on a WAV in tmpfs, warm process, median of "
                    f"{reps} calls; includes WAV parsing, device buffer allocation, kernels and the signatures coming back"}


def live_section(ctx, seconds: float = 5.0, reps: int = 200):
    """The live path (generate_fingerprint_from_samples, fingerprinting.c:81-109: what the ears loop calls once a second):
    latency of one call on a query-sized window of normalised 5512 Hz samples (34 spectral images for 5 s)."""
    rng = np.random.RandomState(3)
    n = int(seconds * 5512.5)
    t = np.arange(n) / 5512.5
    s = 0.35 * np.sin(2 * np.pi * (600 + 350 * np.sin(2 * np.pi * 0.6 * t)) * t) + 0.2 * np.sin(2 * np.pi * 1310 * t) * (t % 0.7 < 0.3)
    s = np.clip(s + 0.04 * rng.standard_normal(n), -1, 1).astype(np.float32)
    for _ in range(5):
        rc, sigs = ctx.index_samples(s)
    lat = []
    for _ in range(reps):
        t0 = time.perf_counter()
        rc, sigs = ctx.index_samples(s)
        lat.append((time.perf_counter() - t0) * 1e3)
    return {"samples": n, "signatures": int(len(sigs)), "p50_ms": float(np.percentile(lat, 50)),
            "p99_ms": float(np.percentile(lat, 99)), "cuda_graph": bool(ctx.live_graph), "calls": reps}


# ------------------------------------------------------------------------------ clocks ----
class ClockSampler:
    """Samples SM clock and throttle reasons in a background thread (NVML, 5 ms period; nvidia-smi as a
    fallback).  mark() .. stop() delimit the timed region; only samples inside it are summarised."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        self.index = index
        self.samples = []          # (t, sm_mhz, reasons_mask)
        self.max_mhz = None
        self._stop = threading.Event()
        self.t0 = None
        self.thread = None
        self.proc = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def loop():
                while not self._stop.is_set():
                    try:
                        mhz = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        try:
                            mask = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                        except Exception:
                            mask = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                        self.samples.append((time.perf_counter(), mhz, mask))
                    except Exception:
                        pass
                    time.sleep(0.005)

            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
        except Exception:
            self._start_smi()
        deadline = time.perf_counter() + 3.0          # make sure sampling is live before timing starts
        while not self.samples and time.perf_counter() < deadline:
            time.sleep(0.01)

    def _start_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)

            def rd():
                bits = [0x8, 0x40, 0x20, 0x4]
                for line in self.proc.stdout:
                    r = [x.strip() for x in line.split(",")]
                    try:
                        mask = sum(b for b, v in zip(bits, r[2:6]) if v.lower().startswith("active"))
                        self.samples.append((time.perf_counter(), float(r[0]), mask))
                        self.max_mhz = float(r[1])
                    except Exception:
                        pass
            self.thread = threading.Thread(target=rd, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def stop(self):
        self._stop.set()
        if self.proc:
            self.proc.terminate()

    def summary(self, t0: float, t1: float) -> dict:
        time.sleep(0.02)
        inside = [(m, k) for (t, m, k) in self.samples if t0 <= t <= t1]
        if not inside:                                   # region shorter than one period: nearest samples
            inside = [(m, k) for (t, m, k) in self.samples[-3:]]
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"], "samples": 0}
        mask = 0
        for _, k in inside:
            mask |= k
        return {"sm_mhz": float(np.median([m for m, _ in inside])), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(n for b, n in self.REASONS.items() if mask & b), "samples": len(inside)}


def _cpulist(text: str) -> list:
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus += list(range(int(a), int(b or a) + 1))
    return cpus


def bind_near_gpu(index: int) -> dict:
    """Pin this process (and thus the first-touch placement of its pinned-host allocations) to the CPUs of the NUMA
    node its GPU hangs off, so that N ranks uploading at once do not cross sockets.  The node comes from the PCI
    device's sysfs entry (numa_node), falling back to NVML's CPU affinity; on a single-node box (or when the platform
    reports nothing: -1) there is nothing to place.  Returns what was found, for the bench line."""
    info = {"numa_nodes": 0, "gpu_numa_node": None, "bound_cpus": None}
    try:
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        info["numa_nodes"] = len(nodes)
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bdf = bus.lower()[-12:]                       # 0000:1b:00.0
        cpus = None
        try:
            node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
            info["gpu_numa_node"] = node
            if node >= 0 and len(nodes) > 1:
                cpus = _cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
        except Exception:
            pass
        if cpus is None and len(nodes) > 1:
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1]
            if len(cpus) >= (os.cpu_count() or 0):
                cpus = None                           # "every CPU": uninformative
        if cpus:
            os.sched_setaffinity(0, cpus)
            info["bound_cpus"] = f"{cpus[0]}-{cpus[-1]} ({len(cpus)})"
    except Exception as exc:
        info["error"] = repr(exc)
    return info


def search_section(ctx, ref_queries: int = 8, n_tracks: int = 96, seconds: float = 30.0):
    """Supplementary search figures (BASELINE.json configs[0] scaled out): n_tracks DISTINCT 30-second synthetic
    tracks are indexed on the GPU into one database; queries are 5-second PCM excerpts of every track at four
    offsets that are not hop-aligned (re-fingerprinted on the GPU, as `mnemophonix search` would)."""
    from mnemophonix_b200 import capi, synth
    pcms = [synth.synth_track(5000 + i, seconds, 1) for i in range(n_tracks)]
    entries, st = ctx.index_pcm_batch(pcms)
    clips, picks = [], []
    for i, p in enumerate(pcms):
        for k in range(4):
            off = int((3.0 + 5.5 * k) * 44100) + 1234 + 97 * k
            clips.append(np.ascontiguousarray(p[off: off + 5 * 44100]))
            picks.append(i)
    queries, st = ctx.index_pcm_batch(clips)
    per = int(np.mean([len(e) for e in entries]))
    t0 = time.perf_counter()
    db = capi.Db(ctx, entries)
    t_build = time.perf_counter() - t0
    prepared = capi.Db.prepare(queries)
    db.search(prepared)
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        res, (L, D) = db.search(prepared)
    dt = (time.perf_counter() - t0) / reps
    n_qsig = int(sum(len(q) for q in queries))
    out = {"queries_per_s": len(queries) / dt, "queries": len(queries), "db_entries": len(entries),
           "db_signatures": int(sum(len(e) for e in entries)),
           "lsh_build_s": t_build, "probed_nodes_per_query_signature": L / n_qsig,
           "deep_checks_per_query_signature": D / n_qsig,
           "algorithmic_GBps": (n_qsig * 300 + 4 * L + 100 * D + 16 * len(queries)) / dt / 1e9,
           "excerpts_identified": int(sum(r[0] == e for r, e in zip(res, picks)))}
    try:
        from oracle import ref as refmod
        import ctypes as C
        if refmod.available():
            r = refmod.Ref()
            keep, ents = [], (C.POINTER(refmod.IndexEntry) * len(entries))()
            for i, e in enumerate(entries):
                e = np.ascontiguousarray(e)
                sg = refmod.Signatures(len(e), C.cast(e.ctypes.data, C.POINTER(C.c_uint8 * 100)))
                ie = refmod.IndexEntry(b"t", b"", b"", b"", C.pointer(sg))
                keep.append((e, sg, ie))
                ents[i] = C.pointer(ie)
            idx = refmod.Index(len(entries), C.cast(ents, C.POINTER(C.POINTER(refmod.IndexEntry))))
            lsh = r.lib.create_hash_tables(C.byref(idx))
            t0 = time.perf_counter()
            ref_res = [r.search(np.ascontiguousarray(queries[i]), C.pointer(idx), lsh) for i in range(ref_queries)]
            out["reference_queries_per_s_1thread"] = ref_queries / (time.perf_counter() - t0)
            out["identical_to_reference_on_sample"] = ref_res == [x[0] for x in res[:ref_queries]]
    except Exception as exc:   # the checker is optional here
        out["reference_check_error"] = repr(exc)
    db.close()
    return out


def c4_checker(entries, queries, pick, got, single_score, single_n, sharded_rows, threads):
    """Checker leg of the configs[3] section (rank 0): the sampled queries answered by the unmodified reference
    (oracle/_ref, host threads: its search() is read-only on the index) and their dense per-entry (score, n_matches)
    rows by the oracle (search.c:55-59), against the GPU's answers and rows."""
    import ctypes as C
    from concurrent.futures import ThreadPoolExecutor
    from oracle import ref as refmod
    from oracle.oracle import Oracle
    out = {}
    o = Oracle()
    odb = o.make_db(entries)
    with ThreadPoolExecutor(threads) as ex:
        t0 = time.perf_counter()
        orc = list(ex.map(lambda i: o.search(odb, queries[i]), pick))
        out["oracle_s_per_query_1thread"] = (time.perf_counter() - t0) * threads / len(pick)
    o_rows = np.stack([r[1] for r in orc]).astype(np.int64)           # [sample, n_entries, 2]
    out["dense_rows_single_gpu_equal_oracle"] = bool(np.array_equal(single_score.astype(np.int64), o_rows[:, :, 0])
                                                     and np.array_equal(single_n.astype(np.int64), o_rows[:, :, 1]))
    out["dense_rows_sharded_equal_oracle"] = bool(np.array_equal(np.asarray(sharded_rows).astype(np.int64), o_rows))
    out["oracle_answers_equal"] = [r[0] for r in orc] == [got[i] for i in pick]
    if refmod.available():
        r = refmod.Ref()
        keep, ents = [], (C.POINTER(refmod.IndexEntry) * len(entries))()
        for i, e in enumerate(entries):
            e = np.ascontiguousarray(e)
            sg = refmod.Signatures(len(e), C.cast(e.ctypes.data, C.POINTER(C.c_uint8 * 100)))
            ie = refmod.IndexEntry(b"t", b"", b"", b"", C.pointer(sg))
            keep.append((e, sg, ie))
            ents[i] = C.pointer(ie)
        idx = refmod.Index(len(entries), C.cast(ents, C.POINTER(C.POINTER(refmod.IndexEntry))))
        t0 = time.perf_counter()
        lsh = r.lib.create_hash_tables(C.byref(idx))
        out["reference_lsh_build_s"] = time.perf_counter() - t0
        with ThreadPoolExecutor(threads) as ex:
            t0 = time.perf_counter()
            ref_res = list(ex.map(lambda i: r.search(np.ascontiguousarray(queries[i]), C.pointer(idx), lsh), pick))
            wall = time.perf_counter() - t0
        out["reference_queries_per_s_1thread"] = len(pick) / (wall * threads)
        out["reference_sample"] = int(len(pick))
        out["reference_host_threads"] = threads
        out["identical_to_reference_on_sample"] = ref_res == [got[i] for i in pick]
    return out


def corpus_checker(pcm, sigs):
    """Checker leg of the corpus sections (rank 0): one ring track through the unmodified reference (oracle/_ref) when it
    is built on this box, else through the oracle port; the GPU's signatures of that track must be identical."""
    from oracle import ref as refmod
    if refmod.available():
        rc, want = refmod.Ref().fingerprint_pcm(pcm)
    else:
        from oracle.oracle import Oracle
        rc, want = Oracle().fingerprint_pcm(pcm)
    return rc == 0 and want.shape == sigs.shape and bool(np.array_equal(want, sigs))


# ----------------------------------------------------------------------- reference arm ----
def reference_throughput(seconds_sample: float, channels: int, steps: int, warmup: int, variant: str = "O2"):
    """Times the unmodified reference (oracle/_ref -O2, generate_fingerprint on a WAV file) on the host
    cores: P = cores // 2 (at most 32) concurrent processes, each with the library's own 8 threads per stage
    (spectralimages.c:11), each fingerprinting the sample; this is the configuration that saturated the
    survey's box."""
    from mnemophonix_b200 import synth
    from oracle import ref as refmod
    if not refmod.available(variant):
        return None
    cores = os.cpu_count() or 1
    # ~1/3 of the reference's pipeline is serial, so one process per 2 cores keeps the box busy
    procs = max(1, min(32, cores // 2))
    pcm = make_track(seconds_sample, channels, seed=1)
    tmpdir = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else tempfile.gettempdir()
    path = os.path.join(tmpdir, f"mnemo_bench_{os.getpid()}.wav")
    synth.write_wav(path, pcm)
    worker = (
        "import sys, time; sys.path.insert(0, %r)\n"
        "from oracle.ref import Ref\n"
        "r = Ref(%r)\n"
        "sys.stdout.write('ready\\n'); sys.stdout.flush()\n"
        "for line in sys.stdin:\n"
        "    t = time.perf_counter(); rc, sigs, *_ = r.fingerprint_wav(%r); dt = time.perf_counter() - t\n"
        "    sys.stdout.write('%%d %%d %%.6f\\n' %% (rc, 0 if sigs is None else len(sigs), dt)); sys.stdout.flush()\n"
    ) % (ROOT, variant, path)
    ps = [subprocess.Popen([sys.executable, "-c", worker], stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True)
          for _ in range(procs)]
    try:
        for p in ps:
            assert p.stdout.readline().strip() == "ready"
        times = []
        n_sigs = None
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            for p in ps:
                p.stdin.write("go\n")
                p.stdin.flush()
            for p in ps:
                rc, n, dt = p.stdout.readline().split()
                assert int(rc) == 0
                n_sigs = int(n)
            if it >= warmup:
                times.append(time.perf_counter() - t0)
    finally:
        for p in ps:
            try:
                p.stdin.close()
                p.wait(timeout=5)
            except Exception:
                p.kill()
        if os.path.exists(path):
            os.remove(path)
    ms = 1e3 * float(np.mean(times))
    hours = procs * seconds_sample / 3600.0
    return {"value": hours / (ms / 1e3), "ms_per_step": ms, "cores": cores, "threads": procs * 8, "procs": procs,
            "sample": f"{procs} concurrent processes x one {seconds_sample:.0f} s {channels}-channel excerpt of the "
                      f"workload, reference -{variant} build, generate_fingerprint() on a WAV in tmpfs, tables pre-warmed",
            "n_sigs": n_sigs}


# ------------------------------------------------------------------------------ main ----
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--seconds", type=float, default=7200.0)
    ap.add_argument("--channels", type=int, default=2)
    ap.add_argument("--cpu-sample-seconds", type=float, default=1200.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--search-c4-tracks", type=int, default=10000,
                    help="database size of the configs[3] search section, sharded over the N GPUs (0 = skip)")
    ap.add_argument("--corpus-hours", type=float, default=1000.0,
                    help="length of the configs[4] streaming corpus, dealt to the N GPUs (0 = skip configs[2] and [4])")
    ap.add_argument("--index-only", action="store_true",
                    help="only the headline index measurement (value, e2e, roofline): for profiler launch lists")
    ap.add_argument("--search-c4-ref-queries", type=int, default=64,
                    help="queries of the configs[3] section also answered by the reference / oracle on rank 0")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": f"index one {args.seconds / 3600:g} h synthetic 44.1 kHz 16-bit {args.channels}-channel "
                          f"track per GPU (BASELINE.json configs[1]); PCM {args.seconds * 44100 * 2 * args.channels / 1e9:.2f} GB "
                          "per track > 126 MB L2, so no L2 flush between steps",
              "tracks_per_gpu": 1, "seconds_per_track": args.seconds, "channels": args.channels,
              "parallelism": f"tracks sharded across {args.gpus} GPU(s), no data-path collective",
              "reference_arm_sample": f"--impl reference times {args.cpu_sample_seconds:.0f} s excerpts of this track (cores // 2 "
                                      "concurrent processes x the library's 8 threads), not the whole 2 h file: audio-hours/s "
                                      "does not depend on the length"}

    if args.impl == "reference":
        if rank != 0:
            return
        r = reference_throughput(args.cpu_sample_seconds, args.channels, max(1, args.steps), max(0, min(args.warmup, 1)))
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built"}))
            return
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "reference",
                             "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    import torch
    from mnemophonix_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
    placement = bind_near_gpu(local_rank)
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL announces its version on stdout when the first communicator is created (NCCL_DEBUG=VERSION in this
        # image); stdout must carry exactly one JSON line, so it points at stderr until that has happened
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    stream = torch.cuda.Stream()
    ctx = capi.Context(local_rank, stream.cuda_stream)
    pcm = make_track(args.seconds, args.channels, seed=1 + rank)
    pinned = torch.from_numpy(pcm).pin_memory()
    n_frames = pcm.shape[0]
    track = [(pinned.data_ptr(), n_frames, args.channels)]
    batch = ctx.batch(track)
    out_pinned = torch.empty((max(batch.cap, 1), 100), dtype=torch.uint8).pin_memory()
    hours_per_step = args.seconds / 3600.0

    # ---- device-resident: value ----
    batch.upload()
    batch.sync()
    for _ in range(max(args.warmup, 3)):
        batch.run()
    batch.sync()
    batch.set_timing(True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t_clk0 = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(args.steps):
            batch.run()
        ev1.record(stream)
    barrier()
    clocks = sampler.summary(t_clk0, time.perf_counter())
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    ms_per_step = ms_total / args.steps
    stage_ms, n_timed = batch.get_timing()
    batch.set_timing(False)
    launches_per_run = batch.launches()
    slots_in_use = batch.record_slots_in_use(0)
    value = world * hours_per_step / (ms_per_step / 1e3)

    # ---- end to end through the C ABI with host buffers: e2e ----
    # A streaming indexer keeps two device-side batches in flight, so step i+1's host->device copy
    # overlaps step i's kernels (two streams: copy engine and SMs run concurrently).  Every step still
    # performs its own pinned-host -> device copy of the whole track, the kernels, and the device ->
    # host copy of its signatures; all of it is inside the timed region.
    stream2 = torch.cuda.Stream()
    ctx2 = capi.Context(local_rank, stream2.cuda_stream)
    batch2 = ctx2.batch(track)
    out_pinned2 = torch.empty((max(batch2.cap, 1), 100), dtype=torch.uint8).pin_memory()
    lanes = [(batch, out_pinned), (batch2, out_pinned2)]

    def issue(i):
        b, _ = lanes[i % 2]
        b.upload()
        b.run()

    def collect(i):
        b, o = lanes[i % 2]
        return b.download(out_ptr=o.data_ptr())

    def e2e_loop(steps):
        issue(0)
        for i in range(1, steps):
            issue(i)
            collect(i - 1)
        return collect(steps - 1)

    e2e_loop(3)
    # K steps, timed three times over (barrier + synchronize on both sides of each); the median is reported and all three
    # are listed: the loop is bound by the host -> device copies, which share the host with whatever else runs on the box
    e2e_each = []
    t_e2e0 = time.perf_counter()
    for _ in range(3):
        barrier()
        t_wall = time.perf_counter()
        _, n_sigs, status = e2e_loop(args.steps)
        barrier()
        e2e_each.append(max_over_ranks((time.perf_counter() - t_wall) * 1e3) / args.steps)
    clocks_e2e = sampler.summary(t_e2e0, time.perf_counter())
    sampler.stop()
    e2e_ms = float(np.median(e2e_each))
    e2e_value = world * hours_per_step / (e2e_ms / 1e3)
    h2d = int(pcm.nbytes)
    d2h = int(batch.cap * 100 + 4)
    assert int(status[0]) == 0 and int(n_sigs[0]) > 0
    assert torch.equal(out_pinned[: int(n_sigs[0])], out_pinned2[: int(n_sigs[0])])
    # copy-only ceiling of this run: every rank re-sends its pinned PCM to its GPU, nothing else, all ranks at once
    batch.sync()
    barrier()
    t_c = time.perf_counter()
    for _ in range(max(4, args.steps // 2)):
        batch.upload()
    batch.sync()
    copy_s = max_over_ranks(time.perf_counter() - t_c) / max(4, args.steps // 2)
    barrier()
    # single-buffer latency of one isolated step (upload -> kernels -> download), for reference
    t0 = time.perf_counter()
    issue(0)
    collect(0)
    torch.cuda.synchronize()
    e2e_single_ms = (time.perf_counter() - t0) * 1e3
    batch2.close()
    ctx2.close()

    # ---- roofline of the dominant kernel ----
    dom = max(stage_ms, key=lambda k: stage_ms[k])
    alg_bytes = BYTES_PER_AUDIO_SECOND.get(args.channels, 44100 * 2 * args.channels + 1076.7) * args.seconds
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = alg_bytes / (stage_ms[dom] / 1e3) / 1e9 if stage_ms[dom] > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(dom)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "traffic_source": "profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the "
                                  "committed ncu --set full capture of this workload (not re-measured in this run)",
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": stage_ms[dom],
                "stage_ms": stage_ms,
                "whole_path_frac": (alg_bytes / (ms_per_step / 1e3) / 1e9) / peak,
                "note": "the path is FP32/FP64-issue bound, not HBM bound (DESIGN.md); frac is the contract's "
                        "algorithmic-bytes figure; stage k2_fingerprint = group records + fingerprint kernel"}
    # The roofline that actually binds the dominant kernel (k1_spectral): the FP32 pipe.  The bit-exact radix-2
    # butterfly network costs 2865 FP32-pipe cycles per frame and SM sub-partition (ncu: 1871 FMA-pipe
    # instructions, 994 of them packed FMUL2/FADD2 at two cycles each; tools/fp32x2_probe.cu measures those
    # issue costs), so the kernel cannot run faster than frames * 2865 / (SMs * 4 * SM clock).
    sm_mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0
    n5512 = int(args.seconds * 44100) // 8
    n_frames_track = 1 + (n5512 - 2048) // 64 if n5512 >= 2048 else 0
    sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
    fp32_floor_ms = n_frames_track * 2865.0 / (sm_count * 4 * sm_mhz * 1e6) * 1e3
    roofline["fp32_pipe"] = {"kernel": "k1_spectral", "pipe_cycles_per_frame": 2865, "frames": n_frames_track,
                             "sm_mhz": sm_mhz, "floor_ms": fp32_floor_ms, "kernel_ms": stage_ms.get("k1_spectral"),
                             "frac": (fp32_floor_ms / stage_ms["k1_spectral"]) if stage_ms.get("k1_spectral") else None}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config, "clocks": clocks, "clocks_e2e": clocks_e2e,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms, "ms_per_step_each_repeat": e2e_each, "isolated_step_ms": e2e_single_ms,
                    "h2d_GBps_aggregate": world * h2d / (e2e_ms / 1e3) / 1e9,
                    "h2d_ceiling_GBps_aggregate": world * h2d / copy_s / 1e9,
                    "frac_of_h2d_ceiling": copy_s / (e2e_ms / 1e3),
                    "host_placement": placement,
                    "note": "two batches in flight on two streams; each step copies its PCM from pinned host "
                            "memory and its signatures back; timed on the host clock around all K steps, three times over, median reported; "
                            "h2d_ceiling = the same ranks sending the same pinned buffers with no kernels, same run"},
            # kernels of the timed regions and their warm-ups: `value` (warm-up + K runs) and `e2e` (3 + 3 K + 1 steps), each run
            # launching launches_per_run kernels (counted by the library); the supplementary sections are not included
            "gpu_launches": int(launches_per_run * (max(args.warmup, 3) + args.steps + 3 + 3 * args.steps + 1)),
            "gpu_launches_per_step": int(launches_per_run),
            "roofline": roofline, "signatures_per_track": int(n_sigs[0]), "group_record_slots_in_use": slots_in_use}

    if args.index_only:
        args.search_c4_tracks, args.no_cpu_baseline, args.corpus_hours = 0, True, 0.0
    if world == 1 and rank == 0 and not args.index_only:
        # worst-case content for the group records: the same pass on a saw-tooth crescendo track
        try:
            wpcm = make_worst_content_track(args.seconds, args.channels, seed=1)
            wpin = torch.from_numpy(wpcm).pin_memory()
            batch.upload([(wpin.data_ptr(), n_frames, args.channels)])
            batch.sync()
            for _ in range(3):
                batch.run()
            batch.sync()
            batch.set_timing(True)
            w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                w0.record(stream)
                for _ in range(args.steps):
                    batch.run()
                w1.record(stream)
            torch.cuda.synchronize()
            wms = w0.elapsed_time(w1) / args.steps
            wstage, _ = batch.get_timing()
            batch.set_timing(False)
            _, wn, wst = batch.download(out_ptr=out_pinned.data_ptr())
            line["value_worst_content"] = {"value": hours_per_step / (wms / 1e3), "unit": UNIT, "ms_per_step": wms,
                                           "stage_ms": wstage, "slowdown_vs_value": wms / ms_per_step,
                                           "group_record_slots_in_use": batch.record_slots_in_use(0),
                                           "signatures": int(wn[0]),
                                           "content": "two steady tones + noise under a 10 s saw-tooth crescendo (level doubling every "
                                                      "second): nearly every image has its own window maximum"}
            del wpin
        except Exception as exc:
            line["value_worst_content"] = {"error": repr(exc)}
        # the tolerance mode of K1 (never the default): same pass, same track, and how far its signatures are from the exact ones
        try:
            fstream = torch.cuda.Stream()
            fctx = capi.Context(local_rank, fstream.cuda_stream, fast_spectra=True)
            fbatch = fctx.batch(track)
            fbatch.upload()
            fbatch.sync()
            for _ in range(3):
                fbatch.run()
            fbatch.sync()
            fbatch.set_timing(True)
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(fstream):
                f0.record(fstream)
                for _ in range(args.steps):
                    fbatch.run()
                f1.record(fstream)
            torch.cuda.synchronize()
            fms = f0.elapsed_time(f1) / args.steps
            fstage, _ = fbatch.get_timing()
            fbatch.set_timing(False)
            fsig, fn, _ = fbatch.download()
            batch.upload(track)
            batch.run()
            esig, en, _ = batch.download()
            same = float((fsig[: int(fn[0])] == esig[: int(en[0])]).all(axis=1).mean()) if int(fn[0]) == int(en[0]) else None
            line["fast_spectra"] = {"value": hours_per_step / (fms / 1e3), "unit": UNIT, "ms_per_step": fms, "stage_ms": fstage,
                                    "speedup_vs_value": ms_per_step / fms, "signatures": [int(en[0]), int(fn[0])],
                                    "signatures_identical_to_exact_fraction": same,
                                    "k1_frac_of_hbm": (alg_bytes / (fstage["k1_spectral"] / 1e3) / 1e9) / peak,
                                    "whole_path_frac_of_hbm": (alg_bytes / (fms / 1e3) / 1e9) / peak,
                                    "note": "tolerance mode of the frame -> band-energy kernel (mnemo_create_ex, "
                                            "MNEMO_FAST_SPECTRA): real-input 1024-point four-step FFT with FMA; images within "
                                            "1e-4, raw bits >= 99.9 %, same identified tracks (profiles/r02_fast_gates.json); "
                                            "NOT the default: value / e2e / roofline above are the bit-exact mode"}
            fbatch.close()
            fctx.close()
        except Exception as exc:
            line["fast_spectra"] = {"error": repr(exc)}
        try:
            line["e2e_dropin"] = dropin_section(pcm, args.seconds, local_rank)
        except Exception as exc:
            line["e2e_dropin"] = {"error": repr(exc)}
        try:
            line["live_latency"] = live_section(ctx)
        except Exception as exc:
            line["live_latency"] = {"error": repr(exc)}
        line["search"] = search_section(ctx)
    if args.search_c4_tracks > 0:
        # BASELINE.json configs[3] at every N: the database sharded over the N ranks, NCCL all-gathers + GPU merge,
        # every answer compared with the single-GPU search, a sample with the reference and the oracle's dense rows
        from mnemophonix_b200 import workloads
        try:
            c4 = workloads.c4_search(ctx, stream, torch.device("cuda", local_rank), rank, world, dist,
                                     n_tracks=args.search_c4_tracks, ref_queries=args.search_c4_ref_queries, checker=c4_checker)
        except Exception as exc:   # a supplementary section must never cost the bench line
            c4 = {"error": repr(exc)}
        if rank == 0:
            if isinstance(c4, dict) and "phase_ms" in c4 and c4["phase_ms"].get("probe"):
                # roofline of the search path's dominant kernel (the probe: search_pair_kernel with the pair index, else
                # search_probe_kernel), same definition as `roofline` above: SURVEY.md §8d's algorithmic bytes of the batch
                # (300 + 4 L + 100 D per query signature; every rank meets every query signature of its shard) over the
                # probe phase, CUDA events on the context's stream, slowest layout-independent figure: rank 0's shard
                alg = (c4["query_signatures"] * 300 + 4 * c4["probed_nodes_per_query_signature"] * c4["query_signatures"]
                       + 100 * c4["deep_checks_per_query_signature"] * c4["query_signatures"])
                ach = alg / (c4["phase_ms"]["probe"] / 1e3) / 1e9
                c4["roofline"] = {"bound": "hbm", "kernel": "search probe (k_pairprobe.cu / k_search.cu)", "achieved": ach,
                                  "peak": peak * world, "unit": "GB/s", "frac": ach / (peak * world),
                                  "algorithmic_bytes_per_batch": alg, "kernel_ms": c4["phase_ms"]["probe"],
                                  "traffic": None, "note": "probe phase includes the query upload and the accumulator "
                                  "memsets; ncu per-launch DRAM bytes: profiles/r02_ncu_search_pairs_summary.md"}
            line["search_c4"] = c4
    if args.corpus_hours > 0:
        # BASELINE.json configs[2] (130 four-minute tracks, the DEFCON-scale batch) and configs[4] (1000 h streamed with
        # overlapped uploads) at every N: tracks dealt to the ranks, no collective on the data path
        from mnemophonix_b200 import workloads
        dev = torch.device("cuda", local_rank)
        for key, kw in (("batch_c3", dict(n_tracks=130, per_batch=max(1, -(-130 // (2 * world))), label="configs[2]: ",
                                          passes=8, checker=corpus_checker)),
                        ("stream_c5", dict(n_tracks=int(round(args.corpus_hours * 15)), per_batch=64, label="configs[4]: "))):
            try:
                sec = workloads.stream_index(local_rank, dev, rank, world, dist, seconds=240.0, **kw)
            except Exception as exc:
                sec = {"error": repr(exc)}
            if rank == 0:
                line[key] = sec
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        r = reference_throughput(args.cpu_sample_seconds, args.channels, 1, 0)
        if r is not None:
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "reference",
                                    "sample": r["sample"], "threads": r["threads"]}
            # also the build exactly as the reference's Makefile ships it (no -O flag), on half the sample
            r0 = reference_throughput(args.cpu_sample_seconds / 2, args.channels, 1, 0, variant="O0")
            if r0 is not None:
                line["cpu_baseline"]["value_as_shipped_O0"] = r0["value"]
        else:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
                                    "sample": "oracle/_ref not available on this box"}
    if rank == 0:
        print(json.dumps(line))
    batch.close()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
