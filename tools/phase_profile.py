#!/usr/bin/env python
"""Time the phases of a timestep separately on the GPU (schedule ops, mass balance, record,
end_timestep) with CUDA events, using wsi_sched_range.  Diagnostic only: running phases
separately skips the fused Land/Catchment mass balance, so state drifts from a real run."""
import json
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wsi_b200 import ensemble as E
from wsi_b200.backend import capi, layout as L
from wsi_b200.backend.engine import GpuRunner
from wsi_b200.backend.probes import ProbeSet

R = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
fm = E.build_config2(R=R, T=64)
ps = ProbeSet(fm).add_flows(E.land_out_arcs(fm))
run = GpuRunner(fm, probes=ps.probes, rec_t0=0, rec_T=8, device="cuda:0")
ops = [L.OPS[o] for o in fm.sched_op]
n = len(ops)

def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

res = {}
res["full_step"] = timed(lambda: run.step(10))
res["surfaces_only"] = timed(lambda: run.sched_range(10, 0, 0, capi.RUN_SURFACES))
res["schedule_only"] = timed(lambda: run.sched_range(10, 0, n, capi.RUN_SCHEDULE))
res["schedule+mb"] = timed(lambda: run.sched_range(10, 0, n, capi.RUN_SCHEDULE | capi.RUN_MASS_BALANCE))
for k, name in enumerate(ops):
    res[f"op{k}:{name}"] = timed(lambda k=k: run.sched_range(10, k, k + 1, capi.RUN_SCHEDULE))
res["mass_balance_only"] = timed(lambda: run.sched_range(10, 0, 0, capi.RUN_MASS_BALANCE))
res["record_only"] = timed(lambda: run.sched_range(10, 0, 0, capi.RUN_RECORD))
res["end_timestep_only"] = timed(lambda: run.sched_range(10, 0, 0, capi.RUN_END_TIMESTEP))
res["empty_launch"] = timed(lambda: run.sched_range(10, 0, 0, 0))
print(json.dumps({k: round(v, 4) for k, v in res.items()}, indent=1))
