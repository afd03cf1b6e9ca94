"""Subtree partition of one dendritic river network.

A push-only dendritic network never sends anything upstream, so an *upland* subtree does not
depend on what lies below it: it can run any number of timesteps ahead.  The partition cuts the
network at direct, unbounded River -> River arcs:

* every cut-off subtree becomes part of an upland **shard** (a self-contained flat model: the
  cut arc now ends in a Waste node of its own, which -- like the river it replaces -- accepts
  everything, so the arc's own capacity logic is unchanged);
* the remaining downstream part is the **trunk**: the cut arcs stay in place (same slot order, so
  every receiver sums its in-arcs in the original order) but start at a ``Boundary`` node that
  releases, at the cut-off river's position in the discharge sweep, the VQIP series that river
  parked on the arc (``pending``: its release, recorded every step by the upland shard).

Because the receiver adds exactly the VQIP the original river released, in the original order,
the partitioned run is bit-identical to the unpartitioned one (tests/test_partition.py).

Shards are independent of each other: they spread over GPUs without any exchange; the trunk
needs their boundary series -- one gather of ``[T, n_cut, V]`` doubles (KBs per step), the only
communication of the scheme (``parallel.run_partitioned``).
"""
from __future__ import annotations

import numpy as np

from . import layout as L
from .flatten import FlatModel

UNBOUNDED = 1e15


def csr_single(has, value, n):
    """CSR (ptr, flat) of lists with at most one element: ``value[i]`` where ``has[i]``."""
    ptr = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(has.astype(np.int32), out=ptr[1:])
    flat = value[has].astype(np.int32)
    return ptr, (flat if len(flat) else np.zeros(1, dtype=np.int32))


def _receiver(fm):
    """receiver[n] (node the single out-arc leads to, -1 for outlets) and out_arc[n]."""
    n_nodes = len(fm.node_type)
    n_out = np.diff(fm.out_ptr)
    if n_out.max(initial=0) > 1:
        raise ValueError("partitioning needs at most one out-arc per node (a dendritic network)")
    out_arc = np.full(n_nodes, -1, dtype=np.int64)
    has = n_out == 1
    out_arc[has] = fm.out_arcs[fm.out_ptr[:-1][has]]
    recv = np.full(n_nodes, -1, dtype=np.int64)
    recv[has] = fm.arc_out[out_arc[has]]
    return recv, out_arc


def _downstream_first(recv):
    """Node order in which every receiver precedes its senders (iterative BFS from the outlets)."""
    n = len(recv)
    order = np.argsort(recv, kind="stable")  # senders grouped by receiver
    starts = np.searchsorted(recv[order], np.arange(-1, n + 1))
    out = np.empty(n, dtype=np.int64)
    head = 0
    roots = order[starts[0]:starts[1]]
    out[:len(roots)] = roots
    tail = len(roots)
    while head < tail:
        v = out[head]
        head += 1
        kids = order[starts[v + 1]:starts[v + 2]]
        out[tail:tail + len(kids)] = kids
        tail += len(kids)
    if tail != n:
        raise ValueError("the network has a cycle")
    return out


def assign_parts(fm, n_parts, granularity=4):
    """part[node]: -1 trunk, k >= 0 upland shard k; cut_roots: rivers whose out-arc is cut."""
    recv, out_arc = _receiver(fm)
    n_nodes = len(recv)
    order = _downstream_first(recv)
    size = np.ones(n_nodes, dtype=np.int64)
    for v in order[::-1]:
        if recv[v] >= 0:
            size[recv[v]] += size[v]
    cap = fm.view("arc", "capacity")[0].min(axis=1)
    river, waste = L.NT["River"], L.NT["Waste"]
    tau = max(n_nodes // (n_parts * granularity), 1)
    sub = np.full(n_nodes, -2, dtype=np.int64)  # -1 trunk, >= 0 subtree id
    roots = []
    for v in order:
        r = recv[v]
        if r < 0:
            sub[v] = -1
        elif sub[r] == -1:
            eligible = (fm.node_type[v] == river and fm.node_type[r] == river and cap[out_arc[v]] >= UNBOUNDED
                        and size[v] <= tau)
            if eligible:
                sub[v] = len(roots)
                roots.append(int(v))
            else:
                sub[v] = -1
        else:
            sub[v] = sub[r]
    # largest-first bin packing of the subtrees onto the shards
    load = np.zeros(n_parts, dtype=np.int64)
    shard_of = np.zeros(len(roots), dtype=np.int64)
    for k in sorted(range(len(roots)), key=lambda k: -size[roots[k]]):
        j = int(np.argmin(load))
        shard_of[k] = j
        load[j] += size[roots[k]]
    part = np.where(sub >= 0, shard_of[np.maximum(sub, 0)] if len(roots) else 0, -1)
    return part, np.array(roots, dtype=np.int64), recv, out_arc


def _extract(fm, keep_nodes, keep_arcs, arc_in, arc_out, node_type, node_gidx_new, groups, sched, names,
             catch_inputs=None):
    """Assemble a FlatModel from already remapped pieces (single out-arc per node)."""
    n_nodes, n_arcs = len(node_type), len(arc_in)
    has_out = np.zeros(n_nodes, dtype=bool)
    out_of = np.zeros(n_nodes, dtype=np.int64)
    has_out[arc_in] = True
    out_of[arc_in] = np.arange(n_arcs)
    out_ptr, out_arcs = csr_single(has_out, out_of, n_nodes)
    order = np.argsort(arc_out, kind="stable").astype(np.int32)  # in-arcs keep the arcs' relative order
    in_ptr = np.zeros(n_nodes + 1, dtype=np.int32)
    np.cumsum(np.bincount(arc_out, minlength=n_nodes), out=in_ptr[1:])
    recv_type = np.zeros(n_nodes, dtype=np.int32)
    recv_type[has_out] = node_type[arc_out[out_of[has_out]]]
    has = np.zeros((n_nodes, L.NF), dtype=bool)
    for f, fname in enumerate(L.FILTER_NAMES):
        types = L.FILTERS[fname]
        has[:, f] = has_out if types is None else has_out & np.isin(recv_type, [L.NT[x] for x in types])
    cand_ptr, cand_arcs = csr_single(has.reshape(-1), np.repeat(out_of, L.NF), n_nodes * L.NF)
    sched_op = np.array([s[0] for s in sched] or [L.OP["NOP"]], dtype=np.int32)
    sched_node = np.array([s[1] for s in sched] or [0], dtype=np.int32)
    return FlatModel(
        pollutants=fm.pollutants, additive=fm.additive, non_additive=fm.non_additive,
        float_accuracy=fm.float_accuracy, R=fm.R, T=fm.T, nq=fm.nq,
        node_names=names[0], node_type=node_type.astype(np.int32), node_gidx=node_gidx_new.astype(np.int32),
        arc_names=names[1], arc_in=arc_in.astype(np.int32), arc_out=arc_out.astype(np.int32),
        out_ptr=out_ptr, out_arcs=out_arcs, in_ptr=in_ptr, in_arcs=order if n_arcs else np.zeros(1, np.int32),
        cand_ptr=cand_ptr, cand_arcs=cand_arcs,
        land_surf_ptr=np.zeros(1, dtype=np.int32), land_surf=np.zeros(1, dtype=np.int32), surf_names=[],
        sched_op=sched_op, sched_node=sched_node, groups=groups, n_obj={k: v.shape[1] // fm.R for k, v in groups.items()},
        node_forcing=fm.node_forcing, surf_inputs=fm.surf_inputs,
        catch_inputs=fm.catch_inputs if catch_inputs is None else catch_inputs,
        month_idx=fm.month_idx, dayofyear=fm.dayofyear, doy_adj=fm.doy_adj, dates=fm.dates,
    )


def _cols(fm, group, objs):
    """Columns of ``group`` for objects ``objs`` (all replicas), in that order."""
    arr = fm.groups[group]
    n = fm.n_obj[group]
    if n == 0 or len(objs) == 0:
        return np.zeros((arr.shape[0], 0))
    return np.ascontiguousarray(arr.reshape(arr.shape[0], n, fm.R)[:, objs, :].reshape(arr.shape[0], -1))


def _submodel(fm, sel, recv, out_arc, cut_roots_here, boundary_of, trunk):
    """Flat model of the nodes ``sel`` (bool mask).

    Upland shard (``trunk`` False): the out-arcs of ``cut_roots_here`` are retargeted to a new
    Waste node.  Trunk: arcs whose sender was cut off keep their slot and start at a new Boundary
    node; ``boundary_of[arc] = series id``."""
    river, catch = L.NT["River"], L.NT["Catchment"]
    n_all = len(fm.node_type)
    keep = np.nonzero(sel)[0]
    new_id = np.full(n_all, -1, dtype=np.int64)
    new_id[keep] = np.arange(len(keep))
    if trunk:
        arcs = np.nonzero(sel[fm.arc_out])[0]  # every arc that ends in the trunk, original order
        cut = arcs[~sel[fm.arc_in[arcs]]]
    else:
        arcs = np.nonzero(sel[fm.arc_in])[0]
        cut = out_arc[cut_roots_here]
    n_extra = len(cut)  # trunk: one Boundary node per cut arc; shard: one Waste sink per cut arc (a single
    #                     sink would have one thread walk all of its in-arcs in the final delivery pass)
    n_nodes = len(keep) + n_extra
    node_type = np.empty(n_nodes, dtype=np.int64)
    node_type[:len(keep)] = fm.node_type[keep]
    node_type[len(keep):] = L.NT["Boundary"] if trunk else L.NT["Waste"]
    a_in = new_id[fm.arc_in[arcs]]
    a_out = new_id[fm.arc_out[arcs]]
    pos = {int(a): i for i, a in enumerate(arcs)}
    cut_pos = np.array([pos[int(a)] for a in cut], dtype=np.int64)
    if trunk:
        a_in[cut_pos] = len(keep) + np.arange(len(cut))
    else:
        a_out[cut_pos] = len(keep) + np.arange(len(cut))
    # groups: keep object order
    g_old = fm.node_gidx
    rivers = keep[fm.node_type[keep] == river]
    catchs = keep[fm.node_type[keep] == catch]
    river_objs = g_old[rivers]
    tank_field = fm.plane("river", "tank")
    tank_objs = fm.groups["river"].reshape(-1, fm.n_obj["river"], fm.R)[tank_field, river_objs, 0].astype(np.int64)
    groups = {k: np.zeros((fm.groups[k].shape[0], 0)) for k in fm.groups}
    groups["arc"] = _cols(fm, "arc", arcs)
    groups["river"] = _cols(fm, "river", river_objs)
    groups["tank"] = _cols(fm, "tank", tank_objs)
    groups["catchm"] = _cols(fm, "catchm", g_old[catchs])
    if len(rivers):
        groups["river"].reshape(-1, len(rivers), fm.R)[tank_field] = np.arange(len(rivers))[:, None]
    gidx = np.zeros(n_nodes, dtype=np.int64)
    gidx[new_id[rivers]] = np.arange(len(rivers))
    gidx[new_id[catchs]] = np.arange(len(catchs))
    if np.any(~np.isin(fm.node_type[keep], [river, catch, L.NT["Waste"], L.NT["Node"]])):
        raise ValueError("partitioning supports River / Catchment / junction / Waste nodes")
    catch_inputs = None
    if trunk:
        n_series = fm.catch_inputs.shape[1]
        gidx[len(keep):] = n_series + np.arange(len(cut))  # series filled in by the driver
        catch_inputs = np.concatenate([fm.catch_inputs, np.zeros((fm.T, len(cut), fm.V))], axis=1)
        for k, a in enumerate(cut):
            boundary_of[int(a)] = n_series + k
    # schedule: ops of kept nodes in order; in the trunk a cut-off river's distribute becomes the
    # boundary node's release at the same position
    bnode = {int(fm.arc_in[a]): len(keep) + k for k, a in enumerate(cut)} if trunk else {}
    sched = []
    for o, n in zip(fm.sched_op, fm.sched_node):
        o, n = int(o), int(n)
        if sel[n]:
            sched.append((o, int(new_id[n])))
        elif trunk and o == L.OP["RIVER_DISTRIBUTE"] and n in bnode:
            sched.append((L.OP["BOUNDARY_RELEASE"], bnode[n]))
    extra = [("boundary:%s" if trunk else "cut_sink:%s") % fm.arc_names[int(a)] for a in cut]
    names = ([fm.node_names[int(i)] for i in keep] + extra, [fm.arc_names[int(a)] for a in arcs])
    sub = _extract(fm, keep, arcs, a_in, a_out, node_type, gidx, groups, sched, names, catch_inputs)
    sub.recipe = {"nodes": keep, "arcs": arcs, "cut_arcs_local": cut_pos, "cut_arcs": np.asarray(cut)}
    return sub


def partition_network(fm, n_parts, granularity=4):
    """Split ``fm`` into ``n_parts`` upland shards and a trunk.

    Returns ``{"shards": [FlatModel], "trunk": FlatModel, "boundary_of": {cut arc: series id in
    trunk.catch_inputs}, "part": per-node part id}``.  Every shard's ``recipe`` maps its nodes /
    arcs back to ``fm`` and lists its cut arcs (whose ``pending`` planes are the boundary series).
    """
    if fm.n_obj["land"] or fm.n_obj["sewer"]:
        raise ValueError("partitioning supports River / Catchment / junction / Waste nodes")
    part, roots, recv, out_arc = assign_parts(fm, n_parts, granularity)
    boundary_of = {}
    trunk = _submodel(fm, part == -1, recv, out_arc, None, boundary_of, trunk=True)
    shards = []
    for k in range(n_parts):
        roots_k = roots[part[roots] == k]
        shards.append(_submodel(fm, part == k, recv, out_arc, roots_k, None, trunk=False))
    return {"shards": shards, "trunk": trunk, "boundary_of": boundary_of, "part": part}


def cut_probes(shard):
    """Probes recording the release parked on every cut arc of an upland shard (V planes each)."""
    o = shard.plane("arc", "pending")
    return [(L.GROUP_ID["arc"], o + p, int(a)) for a in shard.recipe["cut_arcs_local"] for p in range(shard.V)]


def fill_boundary_series(trunk, boundary_of, shard, out_probe, t0=0):
    """Write a shard's recorded cut-arc releases ``out_probe[T', n_cut*V, R]`` (R == 1) into the
    trunk's boundary series for timesteps ``t0 .. t0 + T' - 1``."""
    V = trunk.V
    cuts = shard.recipe["cut_arcs"]
    rec = np.asarray(out_probe)[:, :, 0].reshape(out_probe.shape[0], len(cuts), V)
    for k, a in enumerate(cuts):
        trunk.catch_inputs[t0:t0 + rec.shape[0], boundary_of[int(a)], :] = rec[:, k, :]
