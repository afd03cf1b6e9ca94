"""World-size-2 gloo test of the multi-GPU host logic: members shard across ranks with no
data-path collective; only the mass-balance totals / counts / flags are reduced.  The compute
on each rank is the CPU oracle (this file checks sharding + reductions, not kernels)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
sys.path.insert(0, os.environ['WSI_ROOT']); sys.path.insert(0, os.path.join(os.environ['WSI_ROOT'], 'tests'))
import numpy as np, torch, torch.distributed as dist
from oracle import oracle
from wsi_b200 import ensemble as E, parallel as PAR
from wsi_b200.backend.probes import ProbeSet
dist.init_process_group('gloo')
rank, world = dist.get_rank(), dist.get_world_size()
fm = E.build_config2(R=11, T=60)                 # every rank builds the same global ensemble
ps = ProbeSet(fm).add_flows(E.land_out_arcs(fm))
shard = PAR.shard_members(fm, rank, world)
lo, hi = PAR.shard_bounds(fm.R, rank, world)
assert shard.R == hi - lo
run = oracle.run(shard, probes=ps.probes, n_threads=1)
sums = torch.from_numpy(run.mb_sums.copy()); cnt = torch.from_numpy(run.mb_count.copy())
sums, cnt = PAR.reduce_mass_balance(sums, cnt)
flags = torch.from_numpy(np.bitwise_or.reduce(run.node_flags.reshape(-1, shard.R), axis=1).astype(np.int32))
flags = PAR.reduce_flags(flags)
flows = PAR.gather_outputs(torch.from_numpy(run.out_probe.copy()))
if rank == 0:
    full = oracle.run(fm, probes=ps.probes, n_threads=1)
    assert np.array_equal(flows.numpy(), full.out_probe), 'gathered member outputs differ'
    np.testing.assert_allclose(sums.numpy(), full.mb_sums, rtol=1e-12, atol=1e-9)
    assert np.array_equal(cnt.numpy(), full.mb_count)
    want = np.bitwise_or.reduce(full.node_flags.reshape(-1, fm.R), axis=1).astype(np.int32)
    assert np.array_equal(flags.numpy(), want)
    print('DIST_OK')
dist.destroy_process_group()
"""


def test_two_rank_sharding_and_reductions(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, WSI_ROOT=ROOT, OMP_NUM_THREADS="1")
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
         "127.0.0.1", "--master-port", "29541", str(script)],
        capture_output=True, text=True, env=env, timeout=600)
    assert "DIST_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]


def test_shard_bounds_cover_everything():
    from wsi_b200.parallel import shard_bounds

    for n in (1, 7, 8, 100000):
        for w in (1, 2, 4, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
