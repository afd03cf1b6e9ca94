"""bench.py's reference arm: the UNMODIFIED reference (WSIMOD, installed by baseline/install_ref.sh
into baseline/_ref) timed through its own public API -- ``Model.add_nodes`` / ``add_arcs`` /
``Model.run`` (wsimod/orchestration/model.py:642-778, 893-1200) -- on the config-2 workload.

The reference has no intra-model parallelism, so "all host cores" means one process per core over
disjoint catchments (CPU baseline procedure).  Every catchment is built from the very
node / arc dictionaries the GPU arm's template uses (wsi_b200.ensemble.config2_spec: plain data,
none of this repository's classes or kernels are on the timed path), with its own forcing series and
an object-level parameter perturbation.  A *step* is one timestep of every sampled catchment.

Nothing here reads /root/reference at run time.
"""
from __future__ import annotations

import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_DIR = os.path.join(HERE, "_ref")
NODES_PER_MEMBER = 5


def available():
    return os.path.isdir(os.path.join(REF_DIR, "wsimod"))


def _worker(args):
    """Build ``members`` catchments with the reference's classes, run ``warmup`` untimed and ``steps``
    timed timesteps of each through Model.run; returns (members, seconds inside the timed Model.run
    calls, checksum of the outlet flows)."""
    first, members, warmup, steps, seed = args
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF_DIR)
    sys.path.insert(0, ROOT)
    import numpy as np
    from wsimod.core import constants
    from wsimod.orchestration.model import Model

    from wsi_b200 import ensemble as E  # data only: dates, forcing arrays, the spec dictionaries

    constants.set_default_pollutants()
    constants.FLOAT_ACCURACY = 1e-11
    T = warmup + steps
    dates = E.daily_dates(T)
    forcing = E.synthetic_forcing_arrays(dates, members, seed + first)
    rng = np.random.default_rng(seed + first)
    models = []
    for m in range(members):
        scale = float(np.exp(rng.uniform(np.log(0.8), np.log(1.25))))
        area = float(np.exp(rng.uniform(np.log(0.05), np.log(50.0))))
        nodes, arcs = E.config2_spec(dates, forcing[:, m, :], scale=scale, area_scale=area)
        model = Model()
        model.add_nodes(nodes)
        model.add_arcs(arcs)
        models.append(model)
    for model in models:
        if warmup:
            model.run(dates=dates[:warmup], verbose=False, record_all=False)
    t0 = time.perf_counter()
    check = 0.0
    for model in models:
        flows, _, _, _ = model.run(dates=dates[warmup:], verbose=False, record_all=False,
                                    record_arcs=["land-sewer", "land-gw", "land-river", "river-outlet"])
        check += sum(f["flow"] for f in flows if f["arc"] == "river-outlet")
    return members, time.perf_counter() - t0, check


def run(steps, warmup, members_per_core=8, cores=None, seed=20261019):
    """node-timesteps/s of the reference on ``cores`` processes (default: all) and on one."""
    import multiprocessing as mp

    cores = cores or (os.cpu_count() or 1)
    ctx = mp.get_context("spawn")
    jobs = [(k * members_per_core, members_per_core, warmup, steps, seed) for k in range(cores)]
    with ctx.Pool(cores) as pool:
        pool.map(_worker, [(0, 1, 0, 1, seed)] * cores)  # imports / page-in, untimed
        t0 = time.perf_counter()
        res = pool.map(_worker, jobs)
        wall = time.perf_counter() - t0
        single = pool.map(_worker, [(10_000, members_per_core, warmup, steps, seed)])[0]
    members = sum(r[0] for r in res)
    slowest = max(r[1] for r in res)  # processes run side by side: the job takes as long as the slowest
    return {
        "value": members * steps * NODES_PER_MEMBER / slowest,
        "single_core": single[0] * steps * NODES_PER_MEMBER / single[1],
        "cores": cores,
        "members": members,
        "steps": steps,
        "timed_s": slowest,
        "wall_s": wall,
        "checksum": sum(r[2] for r in res),
    }
