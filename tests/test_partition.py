"""Subtree partition of a dendritic network: shards + trunk reproduce the
unpartitioned run bit for bit; the distributed driver is exercised with world size 2 on gloo.
The kernel logic runs through the host build (tests/hostsim) here; tests/test_gpu_parity.py
repeats the comparison on the GPU."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from wsi_b200 import rivernet
from wsi_b200.backend import layout as L
from wsi_b200.backend import partition as P


class HostsimRunner:
    """GpuRunner's interface on the host build of the kernel logic (network engine)."""

    _lib = None

    def __init__(self, fm, probes, T):
        from hostsim.build import build
        from wsi_b200.backend.capi import ModelBuffers, WsiModelDesc

        if HostsimRunner._lib is None:
            lib = C.CDLL(build())
            lib.wsi_hostsim_run.argtypes = [C.POINTER(WsiModelDesc), C.c_int64, C.c_int64, C.c_int]
            lib.wsi_hostsim_run.restype = C.c_int
            HostsimRunner._lib = lib
        self.fm = fm
        self.buf = ModelBuffers(fm, device="numpy", probes=probes, rec_t0=0, rec_T=T, engine="network")

    def run(self, t0, t1):
        assert self._lib.wsi_hostsim_run(C.byref(self.buf.desc), t0, t1, 0) == 0

    def synchronize(self):
        pass

    def out_probe(self):
        return self.buf.array("out_probe")

    def set_catch_series(self, ids, values, t0=0):
        values = np.asarray(values)
        self.buf.array("catch_inputs")[t0:t0 + len(values), list(ids), :] = values

    def tank_storage(self):
        o = self.fm.plane("tank", "storage")
        return self.buf.array("group_tank")[o:o + self.fm.V].reshape(self.fm.V, -1)


def river_tanks(fm, runner, into, node_ids):
    """Scatter the runner's final river tank storages into ``into[V, n_rivers_total]``."""
    rivers = node_ids[fm.node_type[:len(node_ids)] == L.NT["River"]]
    into[:, rivers] = runner.tank_storage()[:, :len(rivers)]


def outlet_probes(fm):
    waste = [n for n, t in enumerate(fm.node_type) if t == L.NT["Waste"] and fm.node_names[n] == "outlet"]
    if not waste:
        return []
    arcs = [a for a in range(len(fm.arc_out)) if fm.arc_out[a] == waste[0]]
    return [p for a in arcs for p in fm.probes_for_arc(a)]


@pytest.mark.parametrize("n_parts,segments,chunk", [(1, 1, None), (3, 1, None), (3, 3, 7), (2, 4, None)])
def test_partitioned_run_is_bit_identical(n_parts, segments, chunk):
    from wsi_b200.parallel import build_partition_runners, run_partitioned

    N, T = 900, 20
    fm = rivernet.build_dendritic(N, T=T, binding_frac=0.05)
    whole = HostsimRunner(fm, outlet_probes(fm), T)
    whole.run(0, T)
    ref = np.zeros((fm.V, len(fm.node_type)))
    river_tanks(fm, whole, ref, np.arange(len(fm.node_type)))

    parts = P.partition_network(fm, n_parts, stem_segments=segments)
    subs = parts["parts"]
    assert len(subs) == n_parts + segments
    sizes = [len(s.recipe["nodes"]) for s in subs]
    assert sum(sizes) == len(fm.node_type)
    assert min(sizes[:n_parts]) > 0.5 * max(sizes[:n_parts]) and sum(sizes[n_parts:]) < 0.4 * len(fm.node_type)
    out = run_partitioned(parts, build_partition_runners(parts, HostsimRunner, T, extra_probes=outlet_probes), T,
                          chunk=chunk)
    got = np.zeros_like(ref)
    for k, r in out.items():
        river_tanks(subs[k], r, got, subs[k].recipe["nodes"])
    assert np.array_equal(got, ref)
    last = out[len(subs) - 1]  # the segment holding the outlet
    n_cut = len(subs[-1].recipe["cut_out"]) * fm.V
    assert np.array_equal(last.out_probe()[:, n_cut:], whole.out_probe())  # outlet VQIP, every step
    assert not whole.buf.array("mb_count").any()
    assert all(not r.buf.array("mb_count").any() for r in out.values())


def _worker(rank, world, port, q):
    import torch.distributed as dist

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from wsi_b200.parallel import build_partition_runners, run_partitioned

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    N, T = 400, 8
    fm = rivernet.build_dendritic(N, T=T)
    parts = P.partition_network(fm, 4, stem_segments=2)
    runners = build_partition_runners(parts, HostsimRunner, T, rank=rank, world=world, extra_probes=outlet_probes)
    out = run_partitioned(parts, runners, T, rank=rank, world=world, chunk=4)
    if rank == 0:
        q.put(np.asarray(out[len(parts["parts"]) - 1].out_probe()).copy())
    dist.barrier()
    dist.destroy_process_group()


def test_partitioned_run_world_size_2_gloo():
    import torch.multiprocessing as mp

    N, T = 400, 8
    fm = rivernet.build_dendritic(N, T=T)
    whole = HostsimRunner(fm, outlet_probes(fm), T)
    whole.run(0, T)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert np.array_equal(got, whole.out_probe())
