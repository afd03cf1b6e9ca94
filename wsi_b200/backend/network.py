"""Host-side plan for the network engine: level-synchronous routing of a dendritic river
network with one thread per node (wsi_b200/csrc/wsi_net.cuh).

Eligible models contain only River, junction Node, Waste and Catchment nodes, use only the
River.calculate_discharge / Catchment.route orchestration entries plus the river discharge
order, and every sender and junction has one out-arc -- then a push reaches exactly one
receiver, all objects it touches belong to that receiver's thread, and the reference's
sequential, upstream-first routing (model.py:720-735, 1006-1007) equals a level-by-level sweep.
"""
from __future__ import annotations

import numpy as np

from . import layout as L

MAX_JUNCTIONS = 3  # kJunctionDepth in wsi_device.cuh


def empty_plan(n_nodes):
    z = np.zeros(1, dtype=np.int32)
    return {"net_item": np.zeros(8, np.int32), "net_rpath": np.zeros(4, np.int32),
            "net_level_ptr": np.zeros(1, np.int32), "net_level_node": z, "net_calc_node": z, "net_final_node": z,
            "net_bnd_node": z,
            "net_path_ptr": np.zeros(n_nodes + 1, np.int32), "net_path_split": np.zeros(max(n_nodes, 1), np.int32),
            "net_path_sender": z,
            "net_path_arc_ptr": np.zeros(1, np.int32), "net_path_arcs": z}


def why_not(fm):
    """Reason the model cannot use the network engine ('' if it can)."""
    ok_types = {L.NT["Node"], L.NT["Waste"], L.NT["River"], L.NT["Catchment"], L.NT["Boundary"]}
    if not set(int(t) for t in fm.node_type) <= ok_types:
        return "node types other than River / Node / Waste / Catchment"
    SWEEP = (L.OP["RIVER_DISTRIBUTE"], L.OP["BOUNDARY_RELEASE"])
    ok_ops = {L.OP["RIVER_CALC"], L.OP["CATCHMENT_ROUTE"], L.OP["NOP"]} | set(SWEEP)
    if not set(int(o) for o in fm.sched_op) <= ok_ops:
        return "orchestration entries other than River.calculate_discharge / Catchment.route"
    n_out = np.diff(fm.out_ptr)
    for n, t in enumerate(fm.node_type):
        if t == L.NT["Waste"]:
            continue
        if n_out[n] > 1:
            return f"node {fm.node_names[n]} has more than one out-arc"
        if t == L.NT["Catchment"] and n_out[n] != 1:
            return f"catchment {fm.node_names[n]} has no out-arc"
        if t == L.NT["Boundary"] and (
                n_out[n] != 1 or fm.node_type[fm.arc_out[fm.out_arcs[fm.out_ptr[n]]]] not in (L.NT["River"], L.NT["Waste"])):
            return f"boundary {fm.node_names[n]} does not feed a River / Waste directly"
        if t == L.NT["Node"]:
            if n_out[n] != 1:
                return f"junction {fm.node_names[n]} has no outlet"
    rivers = set(int(n) for n in np.nonzero(fm.node_type == L.NT["River"])[0])
    if set(int(n) for o, n in zip(fm.sched_op, fm.sched_node) if int(o) == L.OP["RIVER_CALC"]) != rivers:
        return "River.calculate_discharge is not scheduled for every river"
    # order of ops: every calculate_discharge / route entry must precede the discharge sweep
    ops = [int(o) for o in fm.sched_op if int(o) != L.OP["NOP"]]
    first_dist = next((i for i, o in enumerate(ops) if o in SWEEP), len(ops))
    if any(o not in SWEEP for o in ops[first_dist:]):
        return "orchestration entries after the river discharge sweep"
    # a catchment route must not precede a calculate_discharge of the river it feeds differently
    # from the default order (all calculate_discharge first)
    first_route = next((i for i, o in enumerate(ops) if o == L.OP["CATCHMENT_ROUTE"]), len(ops))
    if any(o == L.OP["RIVER_CALC"] for o in ops[first_route:first_dist]):
        return "River.calculate_discharge scheduled after Catchment.route"
    return ""


def plan_network(fm):
    """Levels and incoming paths, or None when the model is not eligible."""
    if why_not(fm):
        return None
    n_nodes = len(fm.node_type)
    out_arc = [int(fm.out_arcs[fm.out_ptr[n]]) if fm.out_ptr[n + 1] > fm.out_ptr[n] else -1 for n in range(n_nodes)]
    ops = [(int(o), int(n)) for o, n in zip(fm.sched_op, fm.sched_node)]
    calc = [n for o, n in ops if o == L.OP["RIVER_CALC"]]
    sweep = (L.OP["RIVER_DISTRIBUTE"], L.OP["BOUNDARY_RELEASE"])  # boundary nodes stand in the sweep's order
    senders = [n for o, n in ops if o == L.OP["CATCHMENT_ROUTE"]] + [n for o, n in ops if o in sweep]
    boundary = [n for o, n in ops if o == L.OP["BOUNDARY_RELEASE"]]
    distributes = set(n for o, n in ops if o == L.OP["RIVER_DISTRIBUTE"])
    JUNCTION = L.NT["Node"]
    paths = [[] for _ in range(n_nodes)]  # receiver -> [(sender, [arcs])]
    feeds = {}  # sender river -> receiver node
    for s in senders:
        a = out_arc[s]
        if a < 0:
            continue
        arcs = [a]
        node = int(fm.arc_out[a])
        guard = 0
        while fm.node_type[node] == JUNCTION:
            a = out_arc[node]
            arcs.append(a)
            node = int(fm.arc_out[a])
            guard += 1
            if guard > MAX_JUNCTIONS:
                return None  # chain deeper than the kernels' kJunctionDepth (or a cycle)
        if fm.node_type[node] == L.NT["Catchment"]:
            return None
        paths[node].append((s, arcs))
        feeds[s] = node
    # upstream depth of the distributing rivers (iterative, senders before receivers)
    depth = {}
    order = [n for o, n in ops if o == L.OP["RIVER_DISTRIBUTE"]]
    pos = {n: k for k, n in enumerate(order)}
    if any(feeds.get(s) in pos and pos[feeds[s]] < pos[s] for s in order):
        return None  # discharge order is not upstream-first: the sequential semantics differ
    # rivers are given upstream-first by river_discharge_order, but do not rely on it
    indeg = {n: 0 for n in order}
    for s in order:
        r = feeds.get(s)
        if r in indeg:
            indeg[r] += 1
    ready = [n for n in order if indeg[n] == 0]
    for n in ready:
        depth[n] = 0
    i = 0
    while i < len(ready):
        s = ready[i]
        i += 1
        r = feeds.get(s)
        if r in indeg:
            depth[r] = max(depth.get(r, 0), depth[s] + 1)
            indeg[r] -= 1
            if indeg[r] == 0:
                ready.append(r)
    if len(ready) != len(order):
        return None  # cycle among rivers
    n_levels = (max(depth.values()) + 1) if depth else 0
    levels = [[] for _ in range(n_levels)]
    for n in order:
        levels[depth[n]].append(n)
    level_ptr = np.zeros(n_levels + 1, dtype=np.int32)
    level_node = []
    for k, lv in enumerate(levels):
        lv.sort(key=lambda n: pos[n])
        level_node += lv
        level_ptr[k + 1] = len(level_node)
    final = [n for n in range(n_nodes)
             if (fm.node_type[n] == L.NT["Waste"] or (fm.node_type[n] == L.NT["River"] and n not in distributes))
             and paths[n]]
    path_ptr = np.zeros(n_nodes + 1, dtype=np.int32)
    path_split = np.zeros(max(n_nodes, 1), dtype=np.int32)
    path_sender, path_arc_ptr, path_arcs = [], [0], []
    for n in range(n_nodes):
        path_split[n] = len(path_sender) + sum(1 for s, _ in paths[n] if fm.node_type[s] == L.NT["Catchment"])
        for s, arcs in paths[n]:  # already in push order: catchments (orchestration order), then rivers
            path_sender.append(s)
            path_arcs += arcs
            path_arc_ptr.append(len(path_arcs))
        path_ptr[n + 1] = len(path_sender)
    # compact records in level order: what one river of a level needs, without index chasing
    river_tank = fm.view("river", "tank")[0, :, 0].astype(np.int64) if fm.n_obj["river"] else np.zeros(0, np.int64)
    n_out = np.diff(fm.out_ptr)
    item_rec = np.zeros((max(len(level_node), 1), 8), dtype=np.int32)
    rpath_rec = []
    for i, n in enumerate(level_node):
        g = int(fm.node_gidx[n])
        begin = len(rpath_rec)
        for k in range(int(path_split[n]), int(path_ptr[n + 1])):
            s_ = path_sender[k]
            rpath_rec.append((path_arcs[path_arc_ptr[k]],
                              int(river_tank[fm.node_gidx[s_]]) if fm.node_type[s_] == L.NT["River"] else 0,
                              1 if path_arc_ptr[k + 1] - path_arc_ptr[k] == 1 else 0, s_))
        direct = 1 if n_out[n] == 1 and fm.node_type[fm.arc_out[out_arc[n]]] in (L.NT["River"], L.NT["Waste"]) else 0
        item_rec[i] = (n, int(river_tank[g]), g, out_arc[n] if n_out[n] == 1 else -1, begin, len(rpath_rec), direct,
                       rpath_rec[begin][0] if len(rpath_rec) > begin else 0)
    i32 = lambda x: np.array(x if len(x) else [0], dtype=np.int32)  # noqa: E731
    return {
        "net_item": item_rec.reshape(-1), "net_rpath": np.array(rpath_rec or [(0, 0, 0, 0)], dtype=np.int32).reshape(-1),
        "net_level_ptr": level_ptr, "net_level_node": i32(level_node), "net_calc_node": i32(calc),
        "net_final_node": i32(final), "net_bnd_node": i32(boundary), "n_net_bnd": len(boundary),
        "net_path_ptr": path_ptr, "net_path_split": path_split,
        "net_path_sender": i32(path_sender),
        "net_path_arc_ptr": np.array(path_arc_ptr, dtype=np.int32), "net_path_arcs": i32(path_arcs),
        "n_net_calc": len(calc), "n_net_final": len(final), "n_levels": n_levels,
    }
