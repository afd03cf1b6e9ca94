"""Two GPUs, NCCL: the output reductions north_star names (per-step system mass-balance totals,
violation counts, status flags, gather of member outputs) on real sharded runs of the CUDA path,
against the unsharded run.  Skipped on a single-GPU box (the world-size-2 gloo test in
test_distributed_cpu.py covers the host logic there)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
sys.path.insert(0, os.environ['WSI_ROOT']); sys.path.insert(0, os.path.join(os.environ['WSI_ROOT'], 'tests'))
import numpy as np, torch, torch.distributed as dist
from wsi_b200 import ensemble as E, parallel as PAR
from wsi_b200.backend.engine import GpuRunner
from wsi_b200.backend.probes import ProbeSet
rank = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(rank)
dev = f'cuda:{rank}'
dist.init_process_group('nccl', device_id=torch.device(dev))
world = dist.get_world_size()
fm = E.build_config2(R=5001, T=48)               # every rank builds the same global ensemble
ps = ProbeSet(fm).add_flows(E.land_out_arcs(fm))
shard = PAR.shard_members(fm, rank, world)
run = GpuRunner(shard, probes=ps.probes, rec_t0=0, rec_T=fm.T, device=dev)
K = 8                                           # reductions batched every K steps, as bench.py does
sums_d = run.buffers.array('mb_sums'); cnt_d = run.buffers.array('mb_count')
for t0 in range(0, fm.T, K):
    run.run(t0, t0 + K)
    PAR.reduce_mass_balance(sums_d[t0:t0 + K], cnt_d[t0:t0 + K])
run.synchronize()
flags = torch.from_numpy(np.bitwise_or.reduce(run.node_flags().reshape(-1, shard.R), axis=1).astype(np.int32)).to(dev)
flags = PAR.reduce_flags(flags)
flows = PAR.gather_outputs(run.buffers.array('out_probe'))
if rank == 0:
    full = GpuRunner(fm, probes=ps.probes, rec_t0=0, rec_T=fm.T, device=dev)
    full.run(0, fm.T); full.synchronize()
    assert np.array_equal(flows.cpu().numpy(), full.out_probe()), 'gathered member outputs differ'
    np.testing.assert_allclose(sums_d.cpu().numpy(), full.mb_sums(), rtol=1e-12, atol=1e-9)
    assert np.array_equal(cnt_d.cpu().numpy(), full.mb_count())
    want = np.bitwise_or.reduce(full.node_flags().reshape(-1, fm.R), axis=1).astype(np.int32)
    assert np.array_equal(flags.cpu().numpy(), want)
    print('DIST_GPU_OK')
dist.barrier()
dist.destroy_process_group()
"""


def test_two_gpu_sharded_run_and_nccl_reductions(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, WSI_ROOT=ROOT)
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
         "127.0.0.1", "--master-port", "29543", str(script)],
        capture_output=True, text=True, env=env, timeout=900)
    assert "DIST_GPU_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
