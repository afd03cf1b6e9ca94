#!/usr/bin/env python
"""BASELINE.json configs[4] on one GPU: ONE fully coupled synthetic city (golden `city`: 3 districts, 23 nodes --
land with gardens, storm and foul sewers with time-area, groundwater, residential demand, reservoir abstraction,
FWTW, WWTW, three rivers) at an HOURLY step for 20 years = 175 320 timesteps, mass balance checked every step.
A single model is latency-bound (two launches per step, replayed from a CUDA graph by wsi_run_steps); the
golden's 480 hours of forcing and calendar are stepped again and again (state carries over), so the figure is
the cost of 175 320 consecutive timesteps of this model, not a 20-year climate.

    python tools/city_20yr_profile.py > profiles/rNN_config5_city.md
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity_util import load_golden  # noqa: E402

from wsi_b200.backend.engine import GpuRunner  # noqa: E402
from wsi_b200.backend.probes import ProbeSet  # noqa: E402

STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else int(365.25 * 20 * 24)
fm, _ = load_golden("city")
ps = ProbeSet(fm).add_flows()
run = GpuRunner(fm, probes=ps.probes, rec_t0=0, rec_T=fm.T)
run.run(0, fm.T)  # warm-up: builds the graph
run.synchronize()
passes = -(-STEPS // fm.T)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
stream = torch.cuda.current_stream()
w0 = time.perf_counter()
a.record(stream)
done = 0
for _ in range(passes):
    n = min(fm.T, STEPS - done)
    run.run(0, n, stream=stream)
    done += n
b.record(stream)
run.synchronize()
torch.cuda.synchronize()
wall = time.perf_counter() - w0
ms = a.elapsed_time(b)
out = run.out_probe()
assert np.isfinite(out).all()
n = len(fm.node_names)
print("# Config 5: one fully coupled synthetic city, hourly step, 20 years, on one B200\n")
print("`python tools/city_20yr_profile.py` (CUDA events around all steps; wall clock beside it)\n")
print("| nodes | arcs | surfaces | timesteps | device time | wall clock | per step | node-timesteps/s | launches per step |\n|---|---|---|---|---|---|---|---|---|")
print(f"| {n} | {len(fm.arc_names)} | {fm.n_obj['surf']} | {done} | {ms / 1e3:.2f} s | {wall:.2f} s | {ms / done * 1e3:.1f} us | "
      f"{n * done / ms * 1e3:.3e} | {run.launch_count() / (done + fm.T):.2f} |")
print(f"\nEvery arc's VQIP is recorded every step into a {fm.T}-row window; the mass balance of every node is checked every step "
      f"(WSI_B200_FUSED_TAIL_MAX_R = {os.environ.get('WSI_B200_FUSED_TAIL_MAX_R', 'default')}).")
