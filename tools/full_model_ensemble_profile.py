#!/usr/bin/env python
"""Throughput of the schedule kernel's FULL instantiation (pull side, treatment works, demand, distribution):
the reference's complete Oxford demo (18 nodes), the hourly synthetic city and the distribution scenario -- the
golden models of tests/golden -- replicated into ensembles of R identical members and stepped on one GPU.
CUDA events around wsi_run_steps (CUDA-graph replay), after a warm-up pass; no recording beyond the outlet arcs.

    python tools/full_model_ensemble_profile.py [R] > profiles/rNN_full_model_ensembles.md
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity_util import load_golden  # noqa: E402

from wsi_b200.backend.engine import GpuRunner  # noqa: E402
from wsi_b200.backend.probes import ProbeSet  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000
print(f"# Ensembles of the full-featured golden models on one B200 (R = {R} identical members)\n")
print("`python tools/full_model_ensemble_profile.py` (CUDA events, warm-up pass first, CUDA-graph replay)\n")
print("| model | nodes | arcs | surfaces | steps | ms / step | node-timesteps/s |\n|---|---|---|---|---|---|---|")
for name, T in (("oxford", 200), ("city", 96), ("distribution", 200), ("fixture3607_full", 100)):
    fm1, _ = load_golden(name)
    T = min(T, fm1.T)
    fm = fm1.replicate(R)
    ps = ProbeSet(fm).add_flows([len(fm.arc_names) - 1])
    run = GpuRunner(fm, probes=ps.probes, rec_t0=0, rec_T=T)
    run.run(0, T)  # warm-up (also builds the graph)
    run.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream = torch.cuda.current_stream()
    a.record(stream)
    run.run(0, T, stream=stream)
    b.record(stream)
    run.synchronize()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / T
    assert np.isfinite(run.out_probe()).all()
    n = len(fm.node_names)
    print(f"| {name} | {n} | {len(fm.arc_names)} | {fm.n_obj['surf']} | {T} | {ms:.3f} | {n * R / ms * 1e3:.3e} |")
    del run, fm
    torch.cuda.empty_cache()
