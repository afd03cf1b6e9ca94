"""Multi-GPU: members / independent catchments shard across ranks with no data-path
collective.  ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) is used
only for the output reductions: per-step system mass-balance totals, violation counts and
status flags."""
from __future__ import annotations

import numpy as np


def shard_bounds(n_total, rank, world):
    """Contiguous block of members owned by ``rank`` (sizes differ by at most one)."""
    base, extra = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def shard_members(fm, rank, world):
    """Sub-ensemble of flat model ``fm`` owned by ``rank``."""
    lo, hi = shard_bounds(fm.R, rank, world)
    return fm.take_replicas(np.arange(lo, hi))


def reduce_mass_balance(mb_sums, mb_count, group=None):
    """All-reduce the per-step system totals ``[T, 3, V]`` (sum) and violation counts ``[T]``
    (sum) across ranks.  Accepts torch tensors on the backend's device (CUDA for NCCL, CPU
    for gloo); returns the reduced tensors."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(mb_sums, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(mb_count, op=dist.ReduceOp.SUM, group=group)
    return mb_sums, mb_count


def reduce_flags(node_flags_any, group=None):
    """Bitwise-OR of per-node status words across ranks (``[n_nodes]`` int32 tensor)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(node_flags_any, op=dist.ReduceOp.BOR, group=group)
    return node_flags_any


def gather_outputs(local, group=None):
    """Concatenate per-rank member outputs ``[..., R_local]`` along the member axis on every
    rank (all_gather; ranks may own different member counts)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    n = torch.tensor([local.shape[-1]], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes)
    pad = torch.zeros(local.shape[:-1] + (mx,), dtype=local.dtype, device=local.device)
    pad[..., : local.shape[-1]] = local
    parts = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[..., :s] for p, s in zip(parts, sizes)], dim=-1)


# ---- one network partitioned across ranks ---------------------------------
def build_partition_runners(parts, make_runner, T, rank=0, world=1, extra_probes=None):
    """Runners of the upland shards owned by ``rank`` (shard ``k`` belongs to rank ``k % world``)
    and, on rank 0, of the trunk.  ``make_runner(fm, probes, T)`` returns an object with
    ``run(t0, t1)``, ``synchronize()``, ``out_probe()`` (``[T, n_probe, R]`` array or tensor) and
    ``set_catch_series(ids, values[T, len(ids), V])`` -- ``GpuRunner`` with the network engine in
    production, a host build of the kernel logic in the CPU tests.  ``extra_probes(fm)`` may add
    probes after the cut-arc ones."""
    from .backend import partition as P

    runners = {}
    for k, shard in enumerate(parts["shards"]):
        if k % world == rank:
            probes = P.cut_probes(shard) + (list(extra_probes(shard)) if extra_probes else [])
            runners[k] = make_runner(shard, probes, T)
    trunk = None
    if rank == 0:
        trunk = make_runner(parts["trunk"], list(extra_probes(parts["trunk"])) if extra_probes else [], T)
    return {"shards": runners, "trunk": trunk}


def run_partitioned(parts, runners, T, rank=0, world=1, group=None, chunk=None):
    """Run a partitioned network (``backend.partition.partition_network``) for timesteps 0..T-1.

    Upland shards need nothing from each other or from the trunk, so they run ahead.  The
    releases they parked on their cut arcs every step (``[T, n_cut * V]`` doubles per shard) are
    broadcast from their owners -- the scheme's only communication -- written into the trunk's
    boundary series, and rank 0 runs the trunk.

    With GPU runners and ``chunk`` set, every model runs on its own CUDA stream in chunks of
    ``chunk`` timesteps: the shards of a rank overlap each other (each of them alone is bound by
    the latency of its chain of levels, not by the machine), and the trunk works on chunk ``c``
    while the shards are already on chunk ``c + 1``."""
    import torch
    import torch.distributed as dist

    shards, boundary_of = parts["shards"], parts["boundary_of"]
    distributed = world > 1 and dist.is_available() and dist.is_initialized()
    mine = runners["shards"]
    trunk = runners["trunk"] if rank == 0 else None
    n_cut = {k: len(s.recipe["cut_arcs"]) for k, s in enumerate(shards)}
    ids = {k: [boundary_of[int(a)] for a in s.recipe["cut_arcs"]] for k, s in enumerate(shards)}
    gpu = chunk is not None and all(hasattr(r, "private_stream") for r in list(mine.values()) + ([trunk] if trunk else []))
    if not gpu:
        for r in mine.values():
            r.run(0, T)
        series = {}
        for k, r in mine.items():
            r.synchronize()
            series[k] = torch.as_tensor(r.out_probe())[:T, :n_cut[k] * shards[k].V, :]
        if distributed:
            nccl = dist.get_backend(group) == "nccl"
            device = torch.device("cuda", torch.cuda.current_device()) if nccl else torch.device("cpu")
            for k, shard in enumerate(shards):
                buf = (series[k].to(device).contiguous() if k in series
                       else torch.zeros((T, n_cut[k] * shard.V, shard.R), dtype=torch.float64, device=device))
                dist.broadcast(buf, src=k % world, group=group)
                series[k] = buf
        if trunk is not None:
            for k, shard in enumerate(shards):
                if n_cut[k]:
                    trunk.set_catch_series(ids[k], series[k][:, :, 0].reshape(T, n_cut[k], shard.V))
            trunk.run(0, T)
            trunk.synchronize()
        return runners

    # ---- pipelined: one stream per model, chunks of timesteps ----
    device = torch.device("cuda", torch.cuda.current_device())
    link = trunk.private_stream() if trunk is not None else torch.cuda.Stream(device)
    for c0 in range(0, T, chunk):
        c1 = min(T, c0 + chunk)
        done = {}
        for k, r in mine.items():
            st = r.private_stream()
            r.run(c0, c1, stream=st)
            done[k] = torch.cuda.Event()
            done[k].record(st)
        with torch.cuda.stream(link):
            for k, shard in enumerate(shards):
                if n_cut[k] == 0:
                    continue
                n = n_cut[k] * shard.V
                if k in mine:
                    link.wait_event(done[k])
                    buf = mine[k].buffers.array("out_probe")[c0:c1, :n, :]
                else:
                    buf = torch.empty((c1 - c0, n, shard.R), dtype=torch.float64, device=device)
                if distributed:
                    buf = buf.contiguous()
                    dist.broadcast(buf, src=k % world, group=group)
                if trunk is not None:
                    trunk.set_catch_series(ids[k], buf[:, :, 0].reshape(c1 - c0, n_cut[k], shard.V), t0=c0)
            if trunk is not None:
                trunk.run(c0, c1, stream=link)
    torch.cuda.synchronize(device)
    return runners
