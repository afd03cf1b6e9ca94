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
  parked on the arc (``pending``: its release, recorded every step by the upland shard);
* the trunk's main stem -- a chain of thousands of dependent levels, the floor of a timestep --
  can itself be cut into **segments**: segment j works on timestep t while segment j + 1, just
  upstream, is already on a later chunk of timesteps (a pipeline along the river).

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


def _submodel(fm, sel):
    """Flat model of the nodes ``sel`` (bool mask) of ``fm``.

    Arcs leaving the part (sender inside, receiver outside) are retargeted to a new Waste sink each;
    arcs entering it keep their slot (receivers still sum their in-arcs in the original order) and
    start at a new Boundary node each, which stands in the discharge sweep where the cut-off river
    stood and releases series ``n_series + k`` of the part's ``catch_inputs`` (filled by the driver)."""
    river, catch = L.NT["River"], L.NT["Catchment"]
    n_all = len(fm.node_type)
    keep = np.nonzero(sel)[0]
    if np.any(~np.isin(fm.node_type[keep], [river, catch, L.NT["Waste"], L.NT["Node"]])):
        raise ValueError("partitioning supports River / Catchment / junction / Waste nodes")
    new_id = np.full(n_all, -1, dtype=np.int64)
    new_id[keep] = np.arange(len(keep))
    inside_in, inside_out = sel[fm.arc_in], sel[fm.arc_out]
    arcs = np.nonzero(inside_in | inside_out)[0]  # original order
    cut_out = np.nonzero(inside_in[arcs] & ~inside_out[arcs])[0]  # positions in `arcs`
    cut_in = np.nonzero(~inside_in[arcs] & inside_out[arcs])[0]
    n_nodes = len(keep) + len(cut_out) + len(cut_in)
    node_type = np.empty(n_nodes, dtype=np.int64)
    node_type[:len(keep)] = fm.node_type[keep]
    sink0 = len(keep)
    bnd0 = sink0 + len(cut_out)
    node_type[sink0:bnd0] = L.NT["Waste"]
    node_type[bnd0:] = L.NT["Boundary"]
    a_in = new_id[fm.arc_in[arcs]]
    a_out = new_id[fm.arc_out[arcs]]
    a_out[cut_out] = sink0 + np.arange(len(cut_out))
    a_in[cut_in] = bnd0 + np.arange(len(cut_in))
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
    n_series = fm.catch_inputs.shape[1]
    gidx[bnd0:] = n_series + np.arange(len(cut_in))
    catch_inputs = fm.catch_inputs
    if len(cut_in):
        catch_inputs = np.concatenate([fm.catch_inputs, np.zeros((fm.T, len(cut_in), fm.V))], axis=1)
    # schedule: ops of kept nodes in order; a cut-off river's distribute becomes the release of the
    # boundary node that took its place on the arc, at the same position of the sweep
    bnode = {int(fm.arc_in[arcs[q]]): bnd0 + k for k, q in enumerate(cut_in)}
    sched = []
    for o, n in zip(fm.sched_op, fm.sched_node):
        o, n = int(o), int(n)
        if sel[n]:
            sched.append((o, int(new_id[n])))
        elif o == L.OP["RIVER_DISTRIBUTE"] and n in bnode:
            sched.append((L.OP["BOUNDARY_RELEASE"], bnode[n]))
    extra = ["cut_sink:%s" % fm.arc_names[int(arcs[q])] for q in cut_out]
    extra += ["boundary:%s" % fm.arc_names[int(arcs[q])] for q in cut_in]
    names = ([fm.node_names[int(i)] for i in keep] + extra, [fm.arc_names[int(a)] for a in arcs])
    sub = _extract(fm, keep, arcs, a_in, a_out, node_type, gidx, groups, sched, names, catch_inputs)
    sub.recipe = {"nodes": keep, "arcs": arcs,
                  "cut_out_local": cut_out, "cut_out": arcs[cut_out],  # arcs whose release this part records
                  "cut_in": arcs[cut_in], "cut_in_series": n_series + np.arange(len(cut_in))}
    return sub


def _stem_segments(fm, part, recv, out_arc, k):
    """Split the trunk (``part == -1``) into ``k`` segments along its main stem: returns seg[node]
    (0 = the segment holding the outlet, k - 1 the most upstream) for trunk nodes, -1 elsewhere."""
    n = len(recv)
    river = L.NT["River"]
    trunk = part == -1
    seg = np.full(n, -1, dtype=np.int64)
    order = _downstream_first(recv)
    if k <= 1:
        seg[trunk] = 0
        return seg
    height = np.zeros(n, dtype=np.int64)  # longest chain of trunk rivers upstream of (and including) v
    best = np.full(n, -1, dtype=np.int64)
    for v in order[::-1]:
        if not trunk[v] or fm.node_type[v] != river:
            continue
        height[v] += 1
        r = recv[v]
        if r >= 0 and trunk[r] and fm.node_type[r] == river and height[v] > height[r]:
            height[r] = height[v]
            best[r] = v
    roots = [v for v in order if trunk[v] and fm.node_type[v] == river and (recv[v] < 0 or fm.node_type[recv[v]] != river)]
    if not roots:
        seg[trunk] = 0
        return seg
    v = max(roots, key=lambda x: height[x])
    stem = []
    while v >= 0:
        stem.append(int(v))
        v = best[v]
    cap = fm.view("arc", "capacity")[0].min(axis=1)
    cuts = {}
    for j in range(1, k):
        i = round(j * len(stem) / k)
        while i < len(stem) and not (cap[out_arc[stem[i]]] >= UNBOUNDED):
            i += 1  # the cut arc must be unbounded
        if 0 < i < len(stem) and stem[i] not in cuts:
            cuts[stem[i]] = len(cuts) + 1
    for v in order:
        if not trunk[v]:
            continue
        if int(v) in cuts:
            seg[v] = cuts[int(v)]
        else:
            r = recv[v]
            seg[v] = seg[r] if (r >= 0 and trunk[r]) else 0
    return seg


def partition_network(fm, n_parts, granularity=4, stem_segments=1):
    """Split ``fm`` into ``n_parts`` upland shards and a trunk of ``stem_segments`` segments.

    Returns ``{"parts": [FlatModel], "n_upland": n_parts, "links": [(producer, consumer, k, series)],
    "part": per-node part index}``.  Parts are ordered so that a part only feeds parts after it:
    upland shards first, then the trunk's segments from the most upstream to the one holding the
    outlet.  ``links``: cut-out arc ``k`` of ``producer`` (its ``recipe["cut_out"][k]``, recorded by
    ``cut_probes``) is boundary series ``series`` of ``consumer``."""
    if fm.n_obj["land"] or fm.n_obj["sewer"]:
        raise ValueError("partitioning supports River / Catchment / junction / Waste nodes")
    part, roots, recv, out_arc = assign_parts(fm, n_parts, granularity)
    seg = _stem_segments(fm, part, recv, out_arc, stem_segments)
    n_seg = int(seg.max()) + 1
    index = np.where(part >= 0, part, n_parts + (n_seg - 1 - seg))  # upstream segments first
    parts = [_submodel(fm, index == k) for k in range(n_parts + n_seg)]
    series_of = [{int(a): int(sid) for a, sid in zip(p.recipe["cut_in"], p.recipe["cut_in_series"])} for p in parts]
    links = []
    for pi, p in enumerate(parts):
        for k, a in enumerate(p.recipe["cut_out"]):
            q = int(index[fm.arc_out[a]])
            assert q > pi, "a part may only feed parts after it"
            links.append((pi, q, k, series_of[q][int(a)]))
    return {"parts": parts, "n_upland": n_parts, "links": links, "part": index}


def cut_probes(sub):
    """Probes recording the release parked on every cut-out arc of a part (V planes each)."""
    o = sub.plane("arc", "pending")
    return [(L.GROUP_ID["arc"], o + p, int(a)) for a in sub.recipe["cut_out_local"] for p in range(sub.V)]
