"""BASELINE config #4 from PCM: index M synthetic 30-second tracks, then search M five-second excerpts against the
resulting database, sharded across the GPUs of one box on node boundaries of the ranking tree (each rank ranks its own
nodes on its GPU; two small NCCL all-gathers on device tensors per query tile; merge on the GPU), with the parity
checks of mnemophonix_b200/workloads.py:c4_search (single-GPU answers, reference answers and oracle dense rows on a
sample).  The same section runs inside bench.py at every N.

    python tools/bench_c4_pcm.py [--tracks 10000] [--seconds 30] [--queries N] [--ref-queries 64]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_c4_pcm.py

Unlike tools/bench_search_c4.py (signature-level random database: almost no bucket collisions) the signatures here come
from audio, so bucket occupancy is what a real database has (L ~ 42 k probed nodes per query signature at this size,
SURVEY.md §8d).  10 k x 30 s of PCM (26 GB) cannot be synthesised by the numpy generator in bench time, so the same
recipe is evaluated with torch on the GPU; the PCM then takes the normal route: device -> pinned host -> C ABI -> kernels.
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mnemophonix_b200 import capi, workloads
import bench     # the reference / oracle checker leg lives in bench.py

ap = argparse.ArgumentParser()
ap.add_argument("--tracks", type=int, default=10000)
ap.add_argument("--seconds", type=float, default=30.0)
ap.add_argument("--ref-queries", type=int, default=64)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--queries", type=int, default=0)         # 0 = one excerpt per track; else only the first N (profiling)
args = ap.parse_args()
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist = None
if world > 1:
    import torch.distributed as dist
    sys.stdout.flush(); saved = os.dup(1); os.dup2(2, 1)      # NCCL announces its version on stdout
    dist.init_process_group("nccl", device_id=dev)
    dist.all_reduce(torch.zeros(1, device=dev)); torch.cuda.synchronize()
    sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
stream = torch.cuda.Stream()
ctx = capi.Context(lr, stream.cuda_stream)
out = workloads.c4_search(ctx, stream, dev, rank, world, dist, n_tracks=args.tracks, seconds=args.seconds, steps=args.steps,
                          ref_queries=args.ref_queries, n_queries=args.queries, checker=bench.c4_checker)
if rank == 0:
    print(json.dumps(out)); sys.stdout.flush()
ctx.close()
if dist is not None:
    dist.destroy_process_group()
