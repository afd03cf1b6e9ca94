#!/usr/bin/env python
"""one dendritic network partitioned by subtree over the GPUs of a box.

    python tools/net_partition_profile.py [N] [T] [parts] [segments]   # one GPU, every part on its own stream
    python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 --master-port P \
        tools/net_partition_profile.py [N] [T] [parts]                  # shard k on rank k % G, trunk on rank 0

Prints the timing of the unpartitioned run (rank 0) and of the partitioned one (max over ranks of
the upland phase, then the trunk), with CUDA events; setup (build, plan, upload) is not timed."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from wsi_b200 import rivernet  # noqa: E402
from wsi_b200.backend import partition as P  # noqa: E402
from wsi_b200.backend.engine import GpuRunner  # noqa: E402
from wsi_b200.parallel import build_partition_runners, run_partitioned  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 32
rank = int(os.environ.get("RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
n_parts = int(sys.argv[3]) if len(sys.argv) > 3 else max(world, 4)
n_seg = int(sys.argv[4]) if len(sys.argv) > 4 else 1
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl")
dev = f"cuda:{torch.cuda.current_device()}"

fm = rivernet.build_dendritic(N, T=T, names=False)


def make(sub, probes, T_):
    return GpuRunner(sub, probes=probes or [(0, 0, 0)], rec_t0=0, rec_T=T_, device=dev, engine="network")


def timed(fn):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return float(dt.item())


res = {"N": N, "T": T, "world": world, "parts": n_parts}
if rank == 0:
    whole = make(fm, [], T)
    whole.run(0, 2)  # warm-up (first launches, page faults); the run below repeats these steps
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    whole.run(2, T)
    e1.record()
    torch.cuda.synchronize()
    res["whole_ms_per_step"] = e0.elapsed_time(e1) / (T - 2)
    del whole
    torch.cuda.empty_cache()
t0 = time.time()
parts = P.partition_network(fm, n_parts, stem_segments=n_seg)
res["partition_s"] = round(time.time() - t0, 1)
res["segments"] = n_seg
res["part_nodes"] = [len(s.recipe["nodes"]) for s in parts["parts"]]
res["cut_arcs"] = len(parts["links"])
runners = build_partition_runners(parts, make, T, rank=rank, world=world)
res["levels"] = {k: r.buffers.net["n_levels"] for k, r in runners.items()}
CHUNK = int(os.environ.get("WSI_CHUNK", 8))
dt = timed(lambda: run_partitioned(parts, runners, T, rank=rank, world=world, chunk=CHUNK or None))
res["chunk"] = CHUNK
res["partitioned_ms_per_step"] = dt * 1e3 / T
res["node_timesteps_per_s"] = (2 * N + 1) * T / dt
if rank == 0:
    print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in res.items()}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
