"""Host->device copy ceiling for subsets of the box's GPUs (diagnosis of the e2e scaling curve).

One process, one pinned 256 MB buffer and one stream per GPU; for every subset all its GPUs copy at once, repeatedly;
reports the aggregate GB/s.  With several NUMA nodes it also prints the GPU x node matrix (buffer first-touched on each
node's CPUs, one GPU at a time), which is what tells where a rank's pinned ring should live.

    python tools/h2d_subsets.py > gpurun_out/h2d_subsets.json
"""
import json, os, time
import torch

N = 256 << 20
REPS = 12
ng = torch.cuda.device_count()
dst = [torch.empty(N, dtype=torch.uint8, device=f"cuda:{g}") for g in range(ng)]
streams = [torch.cuda.Stream(device=g) for g in range(ng)]


def pinned():
    t = torch.empty(N, dtype=torch.uint8).pin_memory()
    t.fill_(3)
    return t


def run(subset, srcs):
    for g in subset:
        with torch.cuda.stream(streams[g]):
            dst[g].copy_(srcs[g], non_blocking=True)
    for g in subset:
        streams[g].synchronize()
    t0 = time.perf_counter()
    for _ in range(REPS):
        for g in subset:
            with torch.cuda.stream(streams[g]):
                dst[g].copy_(srcs[g], non_blocking=True)
    for g in subset:
        streams[g].synchronize()
    return len(subset) * N * REPS / (time.perf_counter() - t0) / 1e9


out = {"gpus": ng, "bytes_per_copy": N, "cpus": os.cpu_count()}
nodes = sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
out["numa_nodes"] = nodes
src = [pinned() for _ in range(ng)]
subsets = [[g] for g in range(ng)]
if ng >= 2:
    subsets += [[0, 1]]
if ng >= 4:
    subsets += [[0, 2], [0, 3], [0, 1, 2, 3], [0, 1, 2], [2, 3]]
if ng >= 8:
    subsets += [[0, 4], [0, 7], [4, 5, 6, 7], [0, 1, 4, 5], [0, 2, 4, 6], [0, 1, 2, 3, 4, 5], list(range(8))]
out["subsets_GBps"] = {",".join(map(str, s)): round(run(s, src), 2) for s in subsets}
if len(nodes) > 1:
    all_cpus = sorted(os.sched_getaffinity(0))
    matrix = {}
    for node in nodes:
        txt = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        cpus = []
        for part in txt.split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, cpus)
        buf = pinned()                      # first touch on this node
        matrix[str(node)] = [round(run([g], {g: buf}), 2) for g in range(ng)]
        del buf
    os.sched_setaffinity(0, all_cpus)
    out["gpu_by_numa_node_GBps"] = matrix
try:
    import pynvml
    pynvml.nvmlInit()
    topo = []
    for g in range(ng):
        h = pynvml.nvmlDeviceGetHandleByIndex(g)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        try:
            node = int(open(f"/sys/bus/pci/devices/{bus.lower()[-12:]}/numa_node").read())
        except Exception:
            node = None
        topo.append({"gpu": g, "bus": bus, "numa_node": node})
    out["pci"] = topo
except Exception as exc:
    out["pci_error"] = repr(exc)
print(json.dumps(out))
