"""GPU, N >= 2 (skipped on a single-GPU box): one process per GPU over NCCL.
Index: tracks sharded with no data-path collective.  Search: database sharded by entry ranges, one
all-gather of candidate records over NVLink, identical results to the single-GPU search."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from mnemophonix_b200 import capi, dist as mdist, synth
    ctx = capi.Context(rank)
    pcms = [synth.synth_track(6000 + i, 14.0 + 2 * (i % 3), 1 + i % 2) for i in range(10)]
    indexed = mdist.index_sharded(ctx, pcms, rank, world, gather_to=None)
    # every rank needs the whole database for the search shards: exchange on the host (test only)
    parts = [None] * world
    dist.all_gather_object(parts, indexed)
    merged = {}
    for p in parts:
        merged.update(p)
    sigs = [merged[i][0] for i in range(len(pcms))]
    clips = [np.ascontiguousarray(p[5 * 44100 + 321: 10 * 44100 + 321]) for p in pcms]
    qs, _ = ctx.index_pcm_batch(clips)
    res, db = mdist.search_sharded(ctx, sigs, qs + [sigs[7]], rank, world, device=torch.device("cuda", rank))   # record form
    # the production protocol: shards on node boundaries of the ranking tree, ranked on their GPUs, two small
    # all-gathers on device tensors over NCCL, merged on the GPU (dist.RankedShardedSearcher); queries in tiles of 4
    cuts, _ = capi.tree_cuts(len(sigs), world)
    total = sum(len(x) for x in sigs)
    stream = torch.cuda.Stream()
    ctx_s = capi.Context(rank, stream.cuda_stream)
    ranked = mdist.RankedShardedSearcher(ctx_s, sigs[int(cuts[rank]):int(cuts[rank + 1])], total // 2, len(sigs), rank, world,
                                         device=torch.device("cuda", rank), stream=stream, tile_queries=4)
    res_ranked = ranked.search(qs + [sigs[7]])
    assert ranked.message_bytes > 0
    ranked.close()
    assert res_ranked == [tuple(r) for r in res] or res_ranked == res, (res_ranked, res)
    # database shards x query groups (dist.GridShardedSearcher): 2 x 2 on four GPUs, 1 x N (whole replicas) otherwise
    db_shards = 2 if world % 2 == 0 and world > 2 else 1
    shard, _ = mdist.grid_layout(world, db_shards)[rank]
    gcuts, _ = capi.tree_cuts(len(sigs), db_shards)
    grid = mdist.GridShardedSearcher(ctx_s, sigs[int(gcuts[shard]):int(gcuts[shard + 1])], total // 2, len(sigs), rank, world,
                                     db_shards, device=torch.device("cuda", rank), stream=stream)
    res_grid = grid.search(qs + [sigs[7]])
    grid.close()
    ctx_s.close()
    assert [tuple(r) for r in res_grid.tolist()] == [tuple(r) for r in res], (res_grid, res)
    if rank == 0:
        out.put(([len(s) for s in sigs], res, [np.asarray(s).tobytes() for s in sigs], [np.asarray(q).tobytes() for q in qs]))
    dist.barrier()
    db.close()
    ctx.close()
    dist.destroy_process_group()


def test_sharded_index_and_search_over_nccl(ctx, oracle):
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    world = min(world, 4)
    import torch.multiprocessing as mp
    mpc = mp.get_context("spawn")
    out = mpc.Queue()
    port = _free_port()
    procs = [mpc.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    counts, res, sig_bytes, q_bytes = out.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    sigs = [np.frombuffer(b, np.uint8).reshape(-1, 100) for b in sig_bytes]
    qs = [np.frombuffer(b, np.uint8).reshape(-1, 100) for b in q_bytes]
    from mnemophonix_b200 import capi, synth
    # same tracks indexed on one GPU: identical signatures
    pcms = [synth.synth_track(6000 + i, 14.0 + 2 * (i % 3), 1 + i % 2) for i in range(10)]
    single, _ = ctx.index_pcm_batch(pcms)
    assert all(np.array_equal(a, b) for a, b in zip(single, sigs))
    db = capi.Db(ctx, sigs)
    want, _ = db.search(qs + [sigs[7]])
    db.close()
    assert res == want and [r[0] for r in res] == list(range(10)) + [7]
    odb = oracle.make_db(sigs)
    assert [oracle.search(odb, q)[0] for q in qs] == list(range(10))
