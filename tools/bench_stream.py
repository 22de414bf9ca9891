"""BASELINE configs[2] / configs[4]: bulk index of a synthetic corpus with overlapped host -> device PCM upload
(the measurement itself is mnemophonix_b200.workloads.stream_index; bench.py runs the same sections at every N).

    python tools/bench_stream.py [--hours 1000] [--seconds 240] [--per-batch 64] [--distinct 13] [--passes 1]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_stream.py --hours 1000
    python tools/bench_stream.py --hours 8.667 --per-batch 65 --passes 8        # configs[2]: 130 four-minute tracks
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mnemophonix_b200 import workloads
from bench import bind_near_gpu, corpus_checker

ap = argparse.ArgumentParser()
ap.add_argument("--hours", type=float, default=1000.0)
ap.add_argument("--seconds", type=float, default=240.0)
ap.add_argument("--per-batch", type=int, default=64)
ap.add_argument("--distinct", type=int, default=13)
ap.add_argument("--passes", type=int, default=1)
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
out = workloads.stream_index(local_rank, torch.device("cuda", local_rank), rank, world, dist,
                             n_tracks=int(round(args.hours * 3600.0 / args.seconds)), seconds=args.seconds,
                             per_batch=args.per_batch, distinct=args.distinct, check_every=args.check_every,
                             passes=args.passes, checker=corpus_checker)
if rank == 0:
    print(json.dumps(out))
    assert out["tracks_differing_from_single_track_run"] == 0, "a track's signatures depend on its batch"
if dist is not None:
    dist.destroy_process_group()
