"""Known answers of the reference's unit tests for the demand, distribution and sewer nodes (CPU, host build of
the kernel headers).  The reference calls one method of one node (tests/test_demand.py, test_distribution.py,
test_sewer.py); here the smallest model in which the default orchestration makes that call is built through this
package's Model API, flattened and stepped, and the arcs' VQIPs and the nodes' state are compared with the values
the reference's tests expect.  Pollutants as in those tests: phosphate (additive), temperature."""
import ctypes as C

import numpy as np
import pandas as pd
import pytest


@pytest.fixture(scope="module")
def hostsim():
    from hostsim.build import build
    from wsi_b200.backend.capi import WsiModelDesc

    lib = C.CDLL(build())
    lib.wsi_hostsim_run.argtypes = [C.POINTER(WsiModelDesc), C.c_int64, C.c_int64, C.c_int]
    lib.wsi_hostsim_run.restype = C.c_int
    return lib


class Run:
    def __init__(self, hostsim, nodes, arcs, T=1):
        from wsi_b200.backend import layout as L
        from wsi_b200.backend.capi import ModelBuffers
        from wsi_b200.backend.probes import ProbeSet
        from wsi_b200.core import constants
        from wsi_b200.orchestration.model import Model

        constants.set_simple_pollutants()
        constants.FLOAT_ACCURACY = 1e-6
        try:
            self.dates = [pd.Timestamp("2000-01-01") + pd.Timedelta(days=i) for i in range(T)]
            m = Model()
            m.add_nodes([n(self.dates) if callable(n) else n for n in nodes])
            m.add_arcs([dict({"type_": "Arc", "name": a, "in_port": i, "out_port": o}, **kw) for a, i, o, kw in arcs])
            m.dates = self.dates
            fm = m.flatten()
            ps = ProbeSet(fm).add_flows()
            buf = ModelBuffers(fm, device="numpy", probes=ps.probes, rec_t0=0, rec_T=T, record_mb=False, engine="replica")
            assert hostsim.wsi_hostsim_run(C.byref(buf.desc), 0, T, 0) == 0
            self.fm, self.L = fm, L
            self.flows = ps.split(buf.array("out_probe"))["flows"][..., 0]  # [T, arc, V]
            self.groups = buf.groups_numpy()
            self.order = [0] + [fm.pollutants.index(p) + 1 for p in ("phosphate", "temperature")]
        finally:
            constants.set_default_pollutants()
            constants.FLOAT_ACCURACY = 1e-11

    def arc(self, name, t=0):
        """(volume, phosphate, temperature) that travelled the arc in timestep t."""
        return self.flows[t, self.fm.arc_names.index(name)][self.order]

    def field(self, group, name, obj=0):
        g = self.L.GROUPS[group]
        o, n = g.offset(name, self.fm.V), g.size(name, self.fm.V)
        col = self.groups[group][o:o + n, obj]
        return col[self.order] if n == self.fm.V else col


WASTE = {"type_": "Waste", "name": "waste"}


def test_create_demand(hostsim):  # tests/test_demand.py:38-64 (get_demand / create_demand demand.py:75-131)
    r = Run(hostsim, [{"type_": "Demand", "name": "d", "constant_demand": 10, "pollutant_load": {"phosphate": 0.1, "temperature": 12}}, WASTE],
            [("a", "d", "waste", {})])
    assert np.array_equal(r.arc("a"), [10, 0.1, 12])
    # (the reference's test reads total_demand / total_received before Demand.end_timestep empties them again,
    # demand.py:133-137; after a whole timestep they are zero)
    assert np.array_equal(r.field("demand", "total_demand"), [0, 0, 0])


def test_house_demand(hostsim):  # tests/test_demand.py:66-84 (get_house_demand demand.py:303-324)
    def houses(dates):
        return {"type_": "Demand", "node_type_override": "ResidentialDemand", "name": "d", "pollutant_load": {"phosphate": 0.1, "temperature": 0.0},
                "data_input_dict": {("temperature", dates[0]): 10}, "population": 5, "per_capita": 0.12, "constant_weighting": 0.4,
                "constant_temp": 8}

    # house water is pushed to Sewers only (directions of demand.py:229-248)
    r = Run(hostsim, [houses, {"type_": "Sewer", "name": "s", "capacity": 10}, WASTE], [("a", "d", "s", {}), ("b", "s", "waste", {})])
    assert r.arc("a") == pytest.approx([5 * 0.12, 5 * 0.1, 10 * 0.6 + 8 * 0.4], abs=1e-15)


def test_unlimited_distribution(hostsim):  # tests/test_distribution.py:36-48 (pull_set_unlimited distribution.py:170-187)
    r = Run(hostsim, [{"type_": "UnlimitedDistribution", "name": "u"}, {"type_": "Demand", "name": "d", "constant_demand": 1000}, WASTE],
            [("u-d", "u", "d", {}), ("d-w", "d", "waste", {})], T=2)
    for t in range(2):
        assert np.array_equal(r.arc("u-d", t), [1000, 0, 0])
    assert np.array_equal(r.arc("d-w", 1), [1000, 0, 0])


def test_distribution_leakage(hostsim):  # tests/test_distribution.py:50-68 (decorate_leakage_check / _set distribution.py:14-98)
    r = Run(hostsim, [{"type_": "UnlimitedDistribution", "name": "u1"}, {"type_": "Distribution", "name": "d1", "leakage": 0.2},
                      {"type_": "Groundwater", "name": "g1", "capacity": 10}, {"type_": "Demand", "name": "user", "constant_demand": 5}, WASTE],
            [("a1", "u1", "d1", {}), ("a2", "d1", "g1", {}), ("a3", "d1", "user", {}), ("a4", "user", "waste", {})])
    v2 = 5 / (1 - 0.2)
    assert r.arc("a1")[0] == v2  # what the Distribution pulled upstream: the request enlarged by the leak
    assert r.arc("a2")[0] == v2 * 0.2  # the leak, pushed to the Groundwater
    assert r.arc("a3")[0] == pytest.approx(5, abs=1e-15)  # what the puller receives


def test_sewer_land_push_and_discharge(hostsim):  # tests/test_sewer.py:63-92 (push_set_land sewer.py:162-196, make_discharge :198-247)
    # a Demand's push reaches the Sewer under the tag "Demand", which the Sewer handles with push_set_land
    # (sewer.py:84-86), the handler the reference's test calls directly
    r = Run(hostsim, [{"type_": "Demand", "name": "d", "constant_demand": 9, "pollutant_load": {"phosphate": 2, "temperature": 10}},
                      {"type_": "Sewer", "name": "s", "capacity": 10, "pipe_timearea": {0: 0.3, 1: 0.7}}, WASTE],
            [("in", "d", "s", {}), ("out", "s", "waste", {"capacity": 2})])
    assert np.array_equal(r.arc("in"), [9, 2, 10])
    assert r.arc("out") == pytest.approx([2, 2 * 2 / 9, 10], abs=1e-15)
    # the reference's test stops after make_discharge; Sewer.end_timestep then only moves the queue
    assert r.field("sewer", "storage") == pytest.approx([7, 7 * 2 / 9, 10], abs=1e-15)
