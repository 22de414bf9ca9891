"""BASELINE config #5: streaming index of a long synthetic corpus with overlapped host -> device PCM upload.

    python tools/bench_stream.py [--hours 1000] [--seconds 240] [--per-batch 64] [--distinct 13]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_stream.py --hours 1000

The corpus is `hours` of 4-minute mono tracks (1000 h = 15 000 tracks, 317 GB of PCM: more than the host holds, so
the PCM is cycled from a pinned ring of `distinct` synthetic tracks).  Tracks are dealt round-robin
to the ranks (independent units, no data-path collective); every rank keeps two batches in flight on two streams so
that batch k+1's upload overlaps batch k's kernels and batch k-1's signature download.  Every batch's PCM crosses
PCIe again and every batch's signatures come back to pinned host memory, all inside the timed region.

Self-check (size-independent): every occurrence of ring track r must produce byte-identical signatures whatever
batch and position it lands in; the tool compares each track's CRC-32 with the one its ring track produced when
indexed alone before the timed region (parity of those against the CPU oracle is tests/test_gpu_index.py's job).
"""
import argparse, json, os, sys, time, zlib
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mnemophonix_b200 import capi, synth
from bench import bind_near_gpu

ap = argparse.ArgumentParser()
ap.add_argument("--hours", type=float, default=1000.0)
ap.add_argument("--seconds", type=float, default=240.0)
ap.add_argument("--per-batch", type=int, default=64)
ap.add_argument("--distinct", type=int, default=13)
ap.add_argument("--check-every", type=int, default=1)   # CRC every k-th batch (1 = all of them)
args = ap.parse_args()

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
bind_near_gpu(local_rank)
torch.cuda.set_device(local_rank)
dist = None
if world > 1:
    import torch.distributed as dist
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)                       # NCCL's version banner goes to stderr
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dist.all_reduce(torch.zeros(1, device="cuda"))
    torch.cuda.synchronize()
    sys.stdout.flush()
    os.dup2(saved, 1)
    os.close(saved)

n_total = int(round(args.hours * 3600.0 / args.seconds))
mine = list(range(rank, n_total, world))          # global track ids of this rank
ring = [synth.synth_track(1000 + i, args.seconds, 1) for i in range(args.distinct)]
n_frames = ring[0].shape[0]
pinned = [torch.from_numpy(b).pin_memory() for b in ring]

lanes = []
for k in range(2):
    st = torch.cuda.Stream()
    ctx = capi.Context(local_rank, st.cuda_stream)
    lanes.append({"ctx": ctx, "batches": {}})

# golden CRCs: every ring track indexed on its own
gold_crc, gold_n = [], []
for r in range(args.distinct):
    sigs, status = lanes[0]["ctx"].index_pcm_batch([ring[r]])
    assert status == [0]
    gold_crc.append(zlib.crc32(sigs[0].tobytes()))
    gold_n.append(len(sigs[0]))


def get_batch(lane, n):
    d = lanes[lane]["batches"]
    if n not in d:
        b = lanes[lane]["ctx"].batch([(pinned[0].data_ptr(), n_frames, 1)] * n)
        d[n] = (b, torch.empty((max(b.cap, 1), 100), dtype=torch.uint8).pin_memory())
    return d[n]


def stream(track_ids, check_every):
    """Index the given tracks; returns (signatures, tracks checked, mismatches)."""
    jobs = [track_ids[i:i + args.per_batch] for i in range(0, len(track_ids), args.per_batch)]
    total = checked = bad = 0
    pending = None

    def collect(p):
        nonlocal total, checked, bad
        b, out, ids, j = p
        _, n, status = b.download(out_ptr=out.data_ptr())
        total += int(n.sum())
        if any(int(s) for s in status):
            bad += 1
        if j % check_every == 0:
            host = out.numpy()
            off = 0
            for t, cnt in zip(ids, n):
                cnt = int(cnt)
                r = t % args.distinct
                if cnt != gold_n[r] or zlib.crc32(host[off:off + cnt].tobytes()) != gold_crc[r]:
                    bad += 1
                off += cnt
                checked += 1

    for j, ids in enumerate(jobs):
        b, out = get_batch(j % 2, len(ids))
        b.upload([(pinned[t % args.distinct].data_ptr(), n_frames, 1) for t in ids])
        b.run()
        if pending is not None:
            collect(pending)
        pending = (b, out, ids, j)
    if pending is not None:
        collect(pending)
    return total, checked, bad


def barrier():
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()


stream(mine[: 3 * args.per_batch], 1)       # warm-up: buffers allocated, both lanes used
barrier()
t0 = time.perf_counter()
sigs, checked, bad = stream(mine, args.check_every)
barrier()
dt = time.perf_counter() - t0
if dist is not None:
    t = torch.tensor([dt, float(sigs), float(checked), float(bad), float(len(mine))], dtype=torch.float64, device="cuda")
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dt = float(tmax[0])
    sigs, checked, bad, n_done = int(t[1]), int(t[2]), int(t[3]), int(t[4])
else:
    n_done = len(mine)
if rank == 0:
    hours = n_done * args.seconds / 3600.0
    pcm_bytes = n_done * n_frames * 2
    print(json.dumps({
        "config": f"stream {hours:.0f} h of {args.seconds:.0f} s mono tracks ({n_done} tracks, {pcm_bytes / 1e9:.0f} GB PCM cycled "
                  f"from a pinned ring of {args.distinct}) over {world} GPU(s), batches of {args.per_batch}, two batches in "
                  "flight per GPU (upload overlaps kernels and download)",
        "n_gpus": world, "audio_hours": hours, "seconds": dt, "audio_hours_per_s": hours / dt,
        "h2d_GBps_aggregate": pcm_bytes / dt / 1e9, "signatures": sigs, "d2h_bytes": sigs * 100,
        "tracks_crc_checked": checked, "tracks_differing_from_single_track_run": bad}))
    assert bad == 0, "a track's signatures depend on its batch"
if dist is not None:
    dist.destroy_process_group()
