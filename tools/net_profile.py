#!/usr/bin/env python
"""Time the network engine on the synthetic dendritic network with CUDA
events: full step and the routing / mass-balance / record / end_timestep phases separately.
Diagnostic only (phases run separately leave state that drifts from a real run)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from wsi_b200 import rivernet  # noqa: E402
from wsi_b200.backend import capi  # noqa: E402
from wsi_b200.backend.engine import GpuRunner  # noqa: E402
from wsi_b200.backend.probes import ProbeSet  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
T = 24
t0 = time.time()
fm = rivernet.build_dendritic(N, T=T, names=False)
t_build = time.time() - t0
ps = ProbeSet(fm).add_flows([0])  # outlet arc of river slot 1
t0 = time.time()
run = GpuRunner(fm, probes=ps.probes, rec_t0=0, rec_T=T, device="cuda:0", engine="network")
t_plan = time.time() - t0
n = len(fm.sched_op)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


res = {"N": N, "levels": run.buffers.net["n_levels"], "build_s": round(t_build, 1), "plan_upload_s": round(t_plan, 1)}
# a real run first (state stays physical)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
run.run(0, 4)
e0.record()
run.run(4, T)
e1.record()
torch.cuda.synchronize()
res["ms_per_step"] = e0.elapsed_time(e1) / (T - 4)
res["mb_violations"] = int(run.mb_count().sum())
res["routing_only"] = timed(lambda: run.sched_range(10, 0, n, capi.RUN_SCHEDULE))
res["mass_balance_only"] = timed(lambda: run.sched_range(10, 0, 0, capi.RUN_MASS_BALANCE))
res["record_only"] = timed(lambda: run.sched_range(10, 0, 0, capi.RUN_RECORD))
res["end_timestep_only"] = timed(lambda: run.sched_range(10, 0, 0, capi.RUN_END_TIMESTEP))
nodes = 2 * N + 1
res["node_timesteps_per_s"] = nodes / (res["ms_per_step"] * 1e-3)
res["algorithmic_GBps"] = rivernet.algorithmic_bytes_per_river_step(fm, 0) * N / (res["ms_per_step"] * 1e-3) / 1e9
print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in res.items()}))
mode = run.groups()["arc"][fm.plane("arc", "pending_mode")]
print(json.dumps({"arcs_general_mode": int((mode[:N] != 0).sum()), "river_arcs": N}))
