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
