#!/usr/bin/env python
"""Throughput of SURVEY 8d config 3 (Oxford land + river-side subset ensemble) on one GPU:
python tools/config3_profile.py [members]."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402

from parity_util import load_golden  # noqa: E402
from wsi_b200 import ensemble as E  # noqa: E402
from wsi_b200.backend.engine import GpuRunner  # noqa: E402
from wsi_b200.backend.probes import ProbeSet  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 125_000
template, _ = load_golden("config3")
fm = E.build_config3(R, template)
outlet = fm.arc_names.index("mixer_to_waste")
ps = ProbeSet(fm).add_flows([outlet])
run = GpuRunner(fm, probes=ps.probes, rec_t0=0, rec_T=8, engine="replica")
n_nodes = len(fm.node_names)
run.run(0, 8)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 0
e0.record()
for t0 in range(8, 8 + 25 * 8, 8):  # the record ring holds 8 rows: rebase the window every 8 steps
    run.desc.rec_t0 = t0
    run.run(t0, t0 + 8)
    steps += 8
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(json.dumps({"workload": "config3", "members": R, "nodes_per_member": n_nodes, "ms_per_step": round(ms, 4),
                  "node_timesteps_per_s": R * n_nodes / (ms * 1e-3), "mb_violations": int(run.mb_count().sum()),
                  "flags": int(run.node_flags().any())}))
