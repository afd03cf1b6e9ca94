"""Synthetic dendritic river networks: N ``River`` nodes (a share of them
with ``decays``, i.e. DecayTank stores) draining to one ``Waste`` outlet, every river fed by a
lateral ``Catchment`` source.

``parent(i) ~ U{max(0, i - ceil(2 sqrt(i))), ..., i - 1}`` (``default_rng(20261021)``), slot 0 is
the outlet, which gives a tree of depth O(sqrt(N)).  The reference cannot build such a network
beyond ~990 levels (recursive ``assign_upstream``, model.py:780-807) and needs minutes per
thousand nodes, so the flat model is assembled directly: a handful of river *variants* is built
through the drop-in ``Model`` API and flattened (all derived parameters come from the same code
as for any other model), then their SoA columns are tiled over the network with vectorised
numpy.  ``tests/test_rivernet.py`` checks that a small network assembled this way equals the
same network built node by node through the API.
"""
from __future__ import annotations

import math

import numpy as np

from .backend import layout as L
from .backend.flatten import FlatModel, flatten
from .core import constants
from .ensemble import daily_dates

SEED = 20261021

_DECAYS = [
    None,
    {"ammonia": {"constant": 0.03, "exponent": 1.01}},
    {"bod": {"constant": 0.02, "exponent": 1.02}, "nitrite": {"constant": 0.01, "exponent": 1.0}},
    None,
]


def tree_parents(N, seed=SEED):
    """parent[i] for river slots 1..N (0 = outlet)."""
    rng = np.random.default_rng(seed)
    i = np.arange(1, N + 1)
    lo = np.maximum(0, i - np.ceil(2.0 * np.sqrt(i)).astype(np.int64))
    u = rng.random(N)
    parent = lo + np.floor(u * (i - lo)).astype(np.int64)
    return np.minimum(parent, i - 1).astype(np.int32)  # parent[k] is the parent of slot k + 1


def variant_nodes(dates, temperature, n_variants=8):
    """River node dicts (``Model.add_nodes`` input), one per variant."""
    tdata = {("temperature", d): float(temperature[t]) for t, d in enumerate(dates)}
    rng = np.random.default_rng(SEED + 1)
    nodes = []
    for k in range(n_variants):
        length = float(rng.uniform(1500, 9000))
        width = float(rng.uniform(4, 25))
        v = length * width * 0.3
        init = {"volume": v, "do": 0.008, "org-phosphorus": v * 2e-5, "phosphate": v * 5e-5, "ammonia": v * 1e-4,
                "solids": v * 1e-3, "bod": v * 1e-3, "cod": v * 2e-3, "ph": 7.1, "temperature": 9.0,
                "nitrate": v * 2e-3, "nitrite": v * 1e-5, "org-nitrogen": v * 3e-4}
        node = {"type_": "River", "name": "river%d" % k, "length": length, "width": width,
                "data_input_dict": tdata, "initial_storage": init, "damp": float(rng.uniform(0.05, 0.2))}
        if _DECAYS[k % len(_DECAYS)]:
            node["decays"] = dict(_DECAYS[k % len(_DECAYS)])
        nodes.append(node)
    return nodes


def river_variants(dates, temperature, n_variants=8):
    """A model {river_k -> outlet} with one river per variant, through the API."""
    from .orchestration.model import Model

    constants.set_default_pollutants()
    constants.FLOAT_ACCURACY = 1e-11
    rivers = variant_nodes(dates, temperature, n_variants)
    m = Model()
    m.add_nodes([{"type_": "Waste", "name": "outlet"}] + rivers)
    m.add_arcs([{"type_": "Arc", "name": "a%d" % k, "in_port": r["name"], "out_port": "outlet"}
                for k, r in enumerate(rivers)])
    m.dates = dates
    return m


def synthetic_temperature(dates, seed=SEED):
    rng = np.random.default_rng(seed + 3)
    doy = np.array([d.dayofyear for d in dates], dtype=float)
    return 10 + 8 * np.sin(2 * np.pi * (doy - 110) / 365) + rng.normal(0, 2, len(dates))


def catchment_series(T, n_series, V, pollutants, seed=SEED + 2, mean_flow=4e3):
    """[T, n_series, V] lateral inflows: volume + concentrations (Catchment.get_flow's inputs)."""
    rng = np.random.default_rng(seed)
    t = np.arange(T)[:, None]
    phase = rng.uniform(0, 2 * np.pi, n_series)[None, :]
    q = mean_flow * rng.lognormal(0, 0.5, n_series)[None, :] * (1 + 0.6 * np.sin(t / 25.0 + phase))
    q = q * rng.lognormal(0, 0.3, (T, n_series))
    out = np.zeros((T, n_series, V))
    out[:, :, 0] = q
    for j, p in enumerate(pollutants, start=1):
        if p == "temperature":
            out[:, :, j] = 10 + 6 * np.sin(2 * np.pi * (t - 110) / 365) + rng.normal(0, 1, (T, n_series))
        elif p == "ph":
            out[:, :, j] = 7.0 + rng.normal(0, 0.1, (T, n_series))
        elif p == "do":
            out[:, :, j] = 0.009
        else:
            out[:, :, j] = rng.uniform(1e-5, 2e-3, (T, n_series))
    return out


def _csr_single(has, value, n):
    """CSR (ptr, flat) of lists with at most one element: element ``value[i]`` where ``has[i]``."""
    ptr = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(has.astype(np.int32), out=ptr[1:])
    flat = value[has].astype(np.int32)
    return ptr, (flat if len(flat) else np.zeros(1, dtype=np.int32))


def build_dendritic(N, T=365, seed=SEED, n_variants=8, n_series=64, binding_frac=0.0, names=True):
    """Flat model of an N-river dendritic network (2N + 1 nodes, 2N arcs), R = 1."""
    dates = daily_dates(T, "2001-01-01")
    temperature = synthetic_temperature(dates, seed)
    rng = np.random.default_rng(seed + 4)
    fm0 = flatten(river_variants(dates, temperature, n_variants), dates, constants.POLLUTANTS,
                  constants.ADDITIVE_POLLUTANTS, constants.NON_ADDITIVE_POLLUTANTS, constants.FLOAT_ACCURACY)
    V = fm0.V
    n_nodes = 2 * N + 1
    n_arcs = 2 * N
    # slots: 0 outlet | 1..N rivers (lower = further downstream) | N+1..2N catchments
    parent = tree_parents(N, seed)  # of river slot k+1
    node_type = np.empty(n_nodes, dtype=np.int32)
    node_type[0] = L.NT["Waste"]
    node_type[1:N + 1] = L.NT["River"]
    node_type[N + 1:] = L.NT["Catchment"]
    node_gidx = np.zeros(n_nodes, dtype=np.int32)
    node_gidx[1:N + 1] = np.arange(N)
    node_gidx[N + 1:] = np.arange(N)
    # arcs: k in [0, N): river k+1 -> parent; k in [N, 2N): catchment -> river k-N+1
    arc_in = np.concatenate([np.arange(1, N + 1), np.arange(N + 1, 2 * N + 1)]).astype(np.int32)
    arc_out = np.concatenate([parent, np.arange(1, N + 1)]).astype(np.int32)
    has_out = np.ones(n_nodes, dtype=bool)
    has_out[0] = False
    out_of = np.zeros(n_nodes, dtype=np.int64)
    out_of[1:N + 1] = np.arange(N)
    out_of[N + 1:] = np.arange(N, 2 * N)
    out_ptr, out_arcs = _csr_single(has_out, out_of, n_nodes)
    # in-arcs in insertion order: arcs are created river arcs first, then catchment arcs, each
    # by ascending index -- a stable sort of the arcs by receiver reproduces dict insertion order
    order = np.argsort(arc_out, kind="stable").astype(np.int32)
    counts = np.bincount(arc_out, minlength=n_nodes)
    in_ptr = np.zeros(n_nodes + 1, dtype=np.int32)
    np.cumsum(counts, out=in_ptr[1:])
    in_arcs = order
    # candidates per (node, filter): the single out-arc where the receiver's class passes
    recv_type = np.zeros(n_nodes, dtype=np.int32)
    recv_type[has_out] = node_type[arc_out[out_of[has_out]]]
    has = np.zeros((n_nodes, L.NF), dtype=bool)
    for f, fname in enumerate(L.FILTER_NAMES):
        types = L.FILTERS[fname]
        if types is None:
            has[:, f] = has_out
        else:
            has[:, f] = has_out & np.isin(recv_type, [L.NT[x] for x in types])
    cand_ptr, cand_arcs = _csr_single(has.reshape(-1), np.repeat(out_of, L.NF), n_nodes * L.NF)

    # ---- groups: tile the variants' columns ----
    variant = rng.integers(0, n_variants, N)
    river0 = fm0.groups["river"]  # [nf, n_variants]
    tank0 = fm0.groups["tank"]
    river = np.ascontiguousarray(river0[:, variant])
    tank = np.ascontiguousarray(tank0[:, variant])
    river[L.GROUPS["river"].offset("tank", V)] = np.arange(N)
    arc = np.zeros((L.GROUPS["arc"].n_fields(V), n_arcs))
    arc[L.GROUPS["arc"].offset("capacity", V)] = constants.UNBOUNDED_CAPACITY
    arc[L.GROUPS["arc"].offset("preference", V)] = 1.0
    arc[L.GROUPS["arc"].offset("stamp", V)] = -1.0
    if binding_frac > 0:  # a sample of river arcs whose capacity binds at high flow
        pick = rng.random(N) < binding_frac
        arc[L.GROUPS["arc"].offset("capacity", V), :N][pick] = rng.uniform(3e3, 3e4, int(pick.sum()))
    series = rng.integers(0, n_series, N)
    catchm = series.astype(np.float64)[None, :]
    groups = {k: np.zeros((fm0.n_fields(k), 0)) for k in L.GROUPS}
    groups.update(arc=arc, tank=tank, river=river, catchm=catchm)
    n_obj = {k: v.shape[1] for k, v in groups.items()}

    # ---- schedule: calculate_discharge, route, then distribute deepest-first ----
    depth = np.zeros(N + 1, dtype=np.int32)  # arcs to the outlet
    for k in range(1, N + 1):  # parents have lower slots, so one ascending pass suffices
        depth[k] = depth[parent[k - 1]] + 1
    dist_order = np.argsort(-depth[1:], kind="stable").astype(np.int32) + 1
    sched_op = np.concatenate([np.full(N, L.OP["RIVER_CALC"]), np.full(N, L.OP["CATCHMENT_ROUTE"]),
                               np.full(N, L.OP["RIVER_DISTRIBUTE"])]).astype(np.int32)
    sched_node = np.concatenate([np.arange(1, N + 1), np.arange(N + 1, 2 * N + 1), dist_order]).astype(np.int32)

    if names:
        node_names = ["outlet"] + ["river%d" % k for k in range(1, N + 1)] + ["catch%d" % k for k in range(1, N + 1)]
        arc_names = ["r%d" % k for k in range(1, N + 1)] + ["c%d" % k for k in range(1, N + 1)]
    else:
        node_names = _Names(n_nodes, "n")
        arc_names = _Names(n_arcs, "a")
    fm = FlatModel(
        pollutants=fm0.pollutants, additive=fm0.additive, non_additive=fm0.non_additive,
        float_accuracy=fm0.float_accuracy, R=1, T=T, nq=fm0.nq,
        node_names=node_names, node_type=node_type, node_gidx=node_gidx, arc_names=arc_names,
        arc_in=arc_in, arc_out=arc_out, out_ptr=out_ptr, out_arcs=out_arcs, in_ptr=in_ptr, in_arcs=in_arcs,
        cand_ptr=cand_ptr, cand_arcs=cand_arcs,
        land_surf_ptr=np.zeros(1, dtype=np.int32), land_surf=np.zeros(1, dtype=np.int32), surf_names=[],
        sched_op=sched_op, sched_node=sched_node, groups=groups, n_obj=n_obj,
        node_forcing=fm0.node_forcing, surf_inputs=fm0.surf_inputs,
        catch_inputs=catchment_series(T, n_series, V, fm0.pollutants),
        month_idx=fm0.month_idx, dayofyear=fm0.dayofyear, doy_adj=fm0.doy_adj, dates=dates,
    )
    fm.recipe = {"parent": parent, "variant": variant, "series": series}  # what the builder drew (tests)
    return fm


class _Names:
    """Lazy name list for million-node networks (len + indexing, no strings held)."""

    def __init__(self, n, prefix):
        self.n, self.prefix = n, prefix

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(self.n))]
        if i < 0:
            i += self.n
        if not 0 <= i < self.n:
            raise IndexError(i)
        return "%s%d" % (self.prefix, i)

    def __iter__(self):
        return (self[i] for i in range(self.n))

    def index(self, name):
        return int(name[len(self.prefix):])


def network_depth(fm):
    from .backend import network

    plan = network.plan_network(fm)
    return None if plan is None else plan["n_levels"]


def algorithmic_bytes_per_river_step(fm, n_recorded_planes=0):
    """Bytes one River + its Catchment + their two arcs must move per timestep (each state plane
    read once and written once, parameters and forcing read once): the roofline numerator for
    the network workload.  The 21- and 366-day lag rings are summed in full every step because
    the reference sums the Python lists (storage.py:795-805; Neumaier ``sum``), so they count."""
    V = fm.V
    ring = 21 + 366
    tank_state = 2 * V + 2 * L.MAX_DECAY  # storage, storage_, total_decayed/pre_decay
    river_state = 2 * V + 4  # bio_in, bio_out, ring heads/counts
    arc_state = 2 * (1 + 2 * V)  # flow_in, vqip_in, pending of the two arcs
    params = 16 + 5 + 4 * L.MAX_DECAY
    forcing = V + 1
    rw = 2 * (tank_state + river_state + arc_state) + 2  # + the two ring slots written
    return 8 * (rw + ring + params + forcing + n_recorded_planes)
