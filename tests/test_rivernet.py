"""Synthetic dendritic networks: the vectorised flat-model builder equals a
node-by-node build through the drop-in API, and the network engine's level-order routing equals
the oracle's sequential routing on a network of a few thousand nodes."""
import ctypes as C

import numpy as np
import pytest

from wsi_b200 import rivernet
from wsi_b200.backend import layout as L
from wsi_b200.backend import network


def _api_build(fm, T):
    """The same network as ``fm`` (from build_dendritic), node by node through Model.add_nodes."""
    from wsi_b200.backend.flatten import flatten
    from wsi_b200.core import constants
    from wsi_b200.orchestration.model import Model

    N = fm.n_obj["river"]
    dates = fm.dates
    temperature = rivernet.synthetic_temperature(dates)
    variants = rivernet.variant_nodes(dates, temperature)
    rec = fm.recipe
    nodes = [{"type_": "Waste", "name": "outlet"}]
    for k in range(N):
        node = dict(variants[rec["variant"][k]])
        node["name"] = "river%d" % (k + 1)
        nodes.append(node)
    for k in range(N):
        s = rec["series"][k]
        did = {}
        for t, d in enumerate(dates):
            did[("flow", d)] = fm.catch_inputs[t, s, 0]
            for j, p in enumerate(fm.pollutants, start=1):
                did[(p, d)] = fm.catch_inputs[t, s, j]
        nodes.append({"type_": "Catchment", "name": "catch%d" % (k + 1), "data_input_dict": did})
    arcs = [{"type_": "Arc", "name": "r%d" % (k + 1), "in_port": "river%d" % (k + 1),
             "out_port": fm.node_names[rec["parent"][k]]} for k in range(N)]
    arcs += [{"type_": "Arc", "name": "c%d" % (k + 1), "in_port": "catch%d" % (k + 1), "out_port": "river%d" % (k + 1)}
             for k in range(N)]
    m = Model()
    m.add_nodes(nodes)
    m.add_arcs(arcs)
    return flatten(m, dates, constants.POLLUTANTS, constants.ADDITIVE_POLLUTANTS, constants.NON_ADDITIVE_POLLUTANTS,
                   constants.FLOAT_ACCURACY)


def test_vectorised_builder_equals_api_build():
    fm = rivernet.build_dendritic(60, T=20)
    ref = _api_build(fm, 20)
    for name in ("node_type", "node_gidx", "arc_in", "arc_out", "out_ptr", "in_ptr", "in_arcs", "cand_ptr"):
        assert np.array_equal(getattr(fm, name), getattr(ref, name)), name
    n_out = fm.out_ptr[-1]
    assert np.array_equal(fm.out_arcs[:n_out], ref.out_arcs[:n_out])
    n_c = fm.cand_ptr[-1]
    assert np.array_equal(fm.cand_arcs[:n_c], ref.cand_arcs[:n_c])
    assert list(fm.node_names) == list(ref.node_names) and list(fm.arc_names) == list(ref.arc_names)
    for g in ("arc", "tank", "river"):
        assert np.array_equal(fm.groups[g], ref.groups[g]), g
    # the API build numbers its catchment series per node; the builder shares them
    sid = ref.groups["catchm"][0].astype(int)
    assert np.array_equal(ref.catch_inputs[:, sid], fm.catch_inputs[:, fm.groups["catchm"][0].astype(int)])
    assert np.array_equal(fm.node_forcing[:, 0], ref.node_forcing[:, 0])
    # schedule: same ops; the distribute sweep may order ties differently but is upstream-first in both
    for op in ("RIVER_CALC", "CATCHMENT_ROUTE"):
        assert np.array_equal(fm.sched_node[fm.sched_op == L.OP[op]], ref.sched_node[ref.sched_op == L.OP[op]])
    for f in (fm, ref):
        order = f.sched_node[f.sched_op == L.OP["RIVER_DISTRIBUTE"]]
        assert sorted(order) == list(range(1, 61))
        pos = {int(n): i for i, n in enumerate(order)}
        for k in range(60):
            p = int(fm.recipe["parent"][k])
            assert p == 0 or pos[k + 1] < pos[p]


def test_plan_levels_respect_dependencies():
    fm = rivernet.build_dendritic(3000, T=5, binding_frac=0.02)
    plan = network.plan_network(fm)
    assert plan is not None and plan["n_levels"] > 10
    level_of = {}
    lp = plan["net_level_ptr"]
    for l in range(plan["n_levels"]):
        for n in plan["net_level_node"][lp[l]:lp[l + 1]]:
            level_of[int(n)] = l
    assert len(level_of) == 3000
    for k in range(3000):
        p = int(fm.recipe["parent"][k])
        if p:
            assert level_of[k + 1] < level_of[p]
    # every arc is on exactly one path and every path ends at its sender's receiver
    arcs = plan["net_path_arcs"][: plan["net_path_arc_ptr"][-1]]
    assert sorted(arcs) == list(range(6000))
    assert list(plan["net_final_node"][: plan["n_net_final"]]) == [0]


@pytest.fixture(scope="module")
def hostsim():
    from hostsim.build import build
    from wsi_b200.backend.capi import WsiModelDesc

    lib = C.CDLL(build())
    lib.wsi_hostsim_run.argtypes = [C.POINTER(WsiModelDesc), C.c_int64, C.c_int64, C.c_int]
    lib.wsi_hostsim_run.restype = C.c_int
    return lib


def test_network_engine_equals_oracle_on_dendritic_network(hostsim):
    """Kernel logic in level order (host build) == oracle in the reference's sequential order,
    bit for bit, with binding arc capacities in the sample."""
    from oracle import oracle
    from wsi_b200.backend.capi import ModelBuffers
    from wsi_b200.backend.probes import ProbeSet

    fm = rivernet.build_dendritic(1500, T=25, binding_frac=0.05)
    ps = ProbeSet(fm).add_flows()
    buf = ModelBuffers(fm, device="numpy", probes=ps.probes, rec_t0=0, rec_T=fm.T, engine="network")
    assert buf.engine == 1
    assert hostsim.wsi_hostsim_run(C.byref(buf.desc), 0, fm.T, 0) == 0
    ora = oracle.run(fm, probes=ps.probes)
    assert np.array_equal(buf.array("out_probe"), ora.out_probe)
    np.testing.assert_allclose(buf.array("mb_sums"), ora.mb_sums, rtol=1e-12, atol=1e-9)
    assert np.array_equal(buf.array("mb_count"), ora.mb_count)
    cap = fm.view("arc", "capacity")[0, :, 0]
    flow = ps.split(buf.array("out_probe"))["flows"][:, :, 0, 0]
    bound = cap < 1e14
    assert bound.any() and np.isclose(flow[:, bound].max(axis=0), cap[bound]).any(), "no capacity ever binds"
