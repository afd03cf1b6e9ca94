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


def reduce_mass_balance_async(mb_sums, mb_count, group=None):
    """The same reduction without blocking the launching stream: returns the work handles (call
    ``wait()`` on each before reading the tensors); the caller keeps stepping meanwhile."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        return [dist.all_reduce(mb_sums, op=dist.ReduceOp.SUM, group=group, async_op=True),
                dist.all_reduce(mb_count, op=dist.ReduceOp.SUM, group=group, async_op=True)]
    return []


def reduce_flags(node_flags_any, group=None):
    """Bitwise-OR of per-node status words across ranks (``[n_nodes]`` int32 tensor).  NCCL has no
    bitwise reductions, so the words travel as one 0/1 plane per status bit under MAX (gloo would take
    BOR directly; one code path serves both)."""
    import torch
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        n_bits = 8  # WSI_FLAG_* uses 6
        shifts = torch.arange(n_bits, dtype=node_flags_any.dtype, device=node_flags_any.device)
        bits = (node_flags_any.unsqueeze(-1) >> shifts) & 1
        dist.all_reduce(bits, op=dist.ReduceOp.MAX, group=group)
        node_flags_any.copy_((bits << shifts).sum(dim=-1))
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
def _owner(parts, k, world):
    """Upland shard k runs on rank k % world; the trunk's segments stay together on rank 0."""
    return k % world if k < parts["n_upland"] else 0


def build_partition_runners(parts, make_runner, T, rank=0, world=1, extra_probes=None):
    """Runners of the parts owned by ``rank``.  ``make_runner(fm, probes, T)`` returns an object with
    ``run(t0, t1)``, ``synchronize()``, ``out_probe()`` (``[T, n_probe, R]`` array or tensor) and
    ``set_catch_series(ids, values[n, len(ids), V], t0)`` -- ``GpuRunner`` with the network engine in
    production, a host build of the kernel logic in the CPU tests.  ``extra_probes(fm)`` may add
    probes after the cut-arc ones."""
    from .backend import partition as P

    runners = {}
    for k, sub in enumerate(parts["parts"]):
        if _owner(parts, k, world) == rank:
            probes = P.cut_probes(sub) + (list(extra_probes(sub)) if extra_probes else [])
            runners[k] = make_runner(sub, probes, T)
    return runners


def run_partitioned(parts, runners, T, rank=0, world=1, group=None, chunk=None):
    """Run a partitioned network (``backend.partition.partition_network``) for timesteps 0..T-1.

    Parts only feed parts after them (upland shards -> trunk segments, upstream segment ->
    downstream segment), through the releases their rivers parked on the cut arcs every step
    (``V`` doubles per cut arc and step).  Upland shards on other ranks broadcast theirs -- the
    scheme's only communication.

    With GPU runners and ``chunk`` set, every part runs on its own CUDA stream in chunks of
    ``chunk`` timesteps, each chunk waiting only for the same chunk of its producers: the parts
    of a rank overlap (each alone is bound by the latency of its chain of levels, not by the
    machine) and the trunk's segments form a pipeline along the river."""
    import torch
    import torch.distributed as dist

    subs, links, n_up = parts["parts"], parts["links"], parts["n_upland"]
    distributed = world > 1 and dist.is_available() and dist.is_initialized()
    V = subs[0].V
    n_out = [len(s.recipe["cut_out"]) for s in subs]
    feeds = {}  # consumer -> producer -> (cut-out positions in the producer, series ids in the consumer)
    for pi, q, k, sid in links:
        ks, sids = feeds.setdefault(q, {}).setdefault(pi, ([], []))
        ks.append(k)
        sids.append(sid)
    gpu = chunk is not None and all(hasattr(r, "private_stream") for r in runners.values())
    device = torch.device("cuda", torch.cuda.current_device()) if gpu or (
        distributed and dist.get_backend(group) == "nccl") else torch.device("cpu")

    def produced(pi, c0, c1, local):
        """[c1 - c0, n_out, V] releases of part pi (a view of its record buffer, or a receive buffer)."""
        if local:
            r = runners[pi]
            buf = r.buffers.array("out_probe") if gpu else torch.as_tensor(r.out_probe())
            return buf[c0:c1, :n_out[pi] * V, 0].reshape(c1 - c0, n_out[pi], V)
        return torch.empty((c1 - c0, n_out[pi], V), dtype=torch.float64, device=device)

    def feed(q, series, c0):
        for pi, (ks, sids) in feeds.get(q, {}).items():
            runners[q].set_catch_series(sids, series[pi][:, ks, :], t0=c0)

    for c0 in range(0, T, chunk or T):
        c1 = min(T, c0 + (chunk or T))
        done, series = {}, {}
        for k in range(len(subs)):
            mine = k in runners
            if mine:
                r = runners[k]
                if gpu:
                    st = r.private_stream()
                    for pi in feeds.get(k, {}):
                        if pi in done:
                            st.wait_event(done[pi])
                    with torch.cuda.stream(st):
                        feed(k, series, c0)
                    r.run(c0, c1, stream=st)
                    done[k] = torch.cuda.Event()
                    done[k].record(st)
                else:
                    feed(k, series, c0)
                    r.run(c0, c1)
                    r.synchronize()
            if n_out[k] == 0:
                continue
            remote_consumer = distributed and k < n_up  # upland releases go to rank 0's trunk
            if mine or remote_consumer:
                if gpu and mine and remote_consumer:
                    torch.cuda.current_stream().wait_event(done[k])
                buf = produced(k, c0, c1, mine)
                if remote_consumer:
                    buf = buf.to(device).contiguous()
                    dist.broadcast(buf, src=_owner(parts, k, world), group=group)
                    if gpu and rank == 0:
                        done[k] = torch.cuda.Event()
                        done[k].record(torch.cuda.current_stream())
                series[k] = buf
    if gpu:
        torch.cuda.synchronize(device)
    return runners
