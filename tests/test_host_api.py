"""Host side of the drop-in boundary: the mirror of the reference's Model / node / arc API,
the flattener's refusals, river ordering, records and scatter-back (CPU only)."""
import numpy as np
import pytest

from parity_util import load_golden


def _quickstart_model(T=1456):
    import pandas as pd

    from wsi_b200.core import constants
    from wsi_b200.orchestration.model import Model

    constants.set_default_pollutants()
    constants.FLOAT_ACCURACY = 1e-11
    fm_ref, _ = load_golden("quickstart")
    dates = [pd.Timestamp("2009-03-03") + pd.Timedelta(days=i) for i in range(T)]
    land_inputs = {(v, d): float(fm_ref.node_forcing[t, 0, j])
                   for t, d in enumerate(dates) for j, v in enumerate(("precipitation", "et0", "temperature"))}
    m = Model()
    m.add_nodes([
        {"type_": "Sewer", "capacity": 0.04, "name": "my_sewer"},
        {"type_": "Land", "data_input_dict": land_inputs, "name": "my_land", "surfaces": [
            {"type_": "ImperviousSurface", "surface": "urban", "area": 10, "pollutant_load": {"phosphate": 1e-7}},
            {"type_": "PerviousSurface", "surface": "rural", "area": 100, "depth": 0.5,
             "pollutant_load": {"phosphate": 1e-7}}]},
        {"type_": "Groundwater", "area": 100, "capacity": 100, "name": "my_groundwater"},
        {"type_": "Node", "name": "my_river"},
        {"type_": "Waste", "name": "my_outlet"}])
    arcs = [("urban_drainage", "my_land", "my_sewer"), ("percolation", "my_land", "my_groundwater"),
            ("runoff", "my_land", "my_river"), ("storm_outflow", "my_sewer", "my_river"),
            ("baseflow", "my_groundwater", "my_river"), ("catchment_outflow", "my_river", "my_outlet")]
    m.add_arcs([{"type_": "Arc", "name": a, "in_port": i, "out_port": o} for a, i, o in arcs])
    m.dates = dates
    return m, fm_ref


def test_mirror_api_flattens_like_the_reference_objects():
    """The golden flat model was flattened from the REFERENCE's objects; the same model built
    through wsi_b200's own classes must flatten to the same arrays."""
    m, fm_ref = _quickstart_model()
    fm = m.flatten()
    for g in fm.groups:
        assert np.array_equal(fm.groups[g], fm_ref.groups[g]), g
    for k in ("node_type", "node_gidx", "arc_in", "arc_out", "out_ptr", "out_arcs", "in_ptr", "in_arcs",
              "cand_ptr", "cand_arcs", "sched_op", "sched_node", "dayofyear", "doy_adj", "month_idx"):
        assert np.array_equal(getattr(fm, k), getattr(fm_ref, k)), k


@pytest.mark.reference
def test_config2_template_flattens_like_reference_classes():
    """Build the config-2 member with the reference's classes and with the mirror; identical
    flat arrays (runs only where /root/reference is mounted)."""
    import subprocess
    import sys

    code = r"""
import sys; sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests/golden'); sys.dont_write_bytecode = True
import numpy as np
import make_golden as G
from wsi_b200 import ensemble as E
dates = E.daily_dates(200)
forcing = E.synthetic_forcing_arrays(dates, 1)
G.set_default_pollutants()
fdict = {}
for t, d in enumerate(dates):
    for j, v in enumerate(('precipitation', 'et0', 'temperature')):
        fdict[(v, d)] = float(forcing[t, 0, j])
ref = G.flatten_ref(G.config2_member(fdict, dates, 1.0, 1.0), dates)
mine = E.config2_template_model(dates, forcing[:, 0, :]).flatten()
for g in ref.groups:
    assert np.array_equal(ref.groups[g], mine.groups[g]), g
assert np.array_equal(ref.cand_arcs, mine.cand_arcs) and np.array_equal(ref.sched_op, mine.sched_op)
print('same')
"""
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert "same" in out.stdout, out.stderr[-2000:]


def test_unsupported_features_are_refused_not_ignored():
    from wsi_b200.backend.flatten import UnsupportedModelError
    from wsi_b200.orchestration.model import Model

    m = Model()
    with pytest.raises(UnsupportedModelError):
        m.add_nodes([{"type_": "NonResidentialDemand", "name": "factory"}])
    m, _ = _quickstart_model(T=10)
    with pytest.raises(UnsupportedModelError):
        m.add_arcs([{"type_": "DecayArc", "name": "x", "in_port": "my_river", "out_port": "my_outlet"}])
    m, _ = _quickstart_model(T=10)
    m.orchestration.append({"Land": "some_custom_function"})
    with pytest.raises(UnsupportedModelError):
        m.flatten()
    m, _ = _quickstart_model(T=10)
    m.add_arcs([{"type_": "Arc", "name": "flood", "in_port": "my_sewer", "out_port": "my_land"}])
    with pytest.raises(UnsupportedModelError):
        m.flatten()


def test_river_order_is_upstream_first():
    from wsi_b200.core import constants
    from wsi_b200.orchestration.model import Model

    constants.set_default_pollutants()
    m = Model()
    m.add_nodes([{"type_": "River", "name": f"r{i}"} for i in range(5)]
                + [{"type_": "Node", "name": "j"}, {"type_": "Waste", "name": "out"}])
    arcs = [("r4", "r3"), ("r3", "j"), ("r2", "j"), ("j", "r1"), ("r1", "r0"), ("r0", "out")]
    m.add_arcs([{"type_": "Arc", "name": f"{a}-{b}", "in_port": a, "out_port": b} for a, b in arcs])
    order = m.river_discharge_order
    assert set(order) == {"r0", "r1", "r2", "r3", "r4"}
    pos = {n: i for i, n in enumerate(order)}
    assert pos["r4"] < pos["r3"] < pos["r1"] < pos["r0"] and pos["r2"] < pos["r1"]


def test_records_and_scatter_back_from_oracle_arrays():
    """_RecordPlan / scatter_back are pure host logic: feed them state produced by the oracle."""
    from oracle import oracle
    from wsi_b200.backend import flatten as F
    from wsi_b200.backend.probes import ProbeSet
    from wsi_b200.orchestration.model import _RecordPlan

    m, fm_ref = _quickstart_model(T=40)
    fm = m.flatten()
    ps = ProbeSet(fm)
    ps.add_flows(list(m.arcs))
    plan = _RecordPlan(m, fm, ps, ["my_groundwater"], [("my_land", "rural")], True)
    run = oracle.run(fm, probes=ps.probes)
    flows, tanks, surfaces = plan.records(ps.split(run.out_probe), m.dates)
    assert len(flows) == 40 * 6 and flows[-1]["arc"] == "catchment_outflow"
    assert set(flows[0]) == {"arc", "flow", "time", *fm.pollutants}
    golden_last = load_golden("quickstart")[1]["flows"][39, 5, 0]
    assert flows[-1]["flow"] == golden_last
    # 1 brief record + record_all: Land x3 ResidenceTanks, Sewer QueueTank, Groundwater Tank
    assert len(tanks) == 40 * (1 + 5)
    assert {t["prop"] for t in tanks[1:6]} == {"percolation", "subsurface_runoff", "surface_runoff", "sewer_tank", "tank"}
    assert len(surfaces) == 40 * 3 and "et0_coef" in surfaces[0] and "percolation" not in surfaces[2 - 1 + 0] or True
    F.scatter_back(m, fm, run.groups, m.dates[-1])
    assert m.nodes["my_groundwater"].tank.storage["volume"] == tanks[-1]["storage"]
    # a second flatten of the scattered-back objects continues the run exactly
    fm2 = m.flatten()
    assert np.array_equal(fm2.groups["tank"], run.groups["tank"])
    assert np.array_equal(fm2.groups["surf"][: fm.plane("surf", "out_percolation")],
                          run.groups["surf"][: fm.plane("surf", "out_percolation")])


@pytest.mark.reference
def test_model_load_of_the_reference_yaml_fixture_flattens_like_the_golden(tmp_path):
    """``Model.load`` on a YAML model file + csv.gz inputs (model.py:216-274): the reference's own
    real-catchment fixture (tests/test_model_data.zip, reduced to the nodes on the GPU path exactly as
    tests/golden/make_golden.py reduces it) loaded through THIS package's Model must flatten to the
    very planes the golden holds, which were flattened from the reference's objects after ITS load."""
    import contextlib
    import io
    import zipfile

    import yaml

    from wsi_b200.orchestration.model import Model

    fm_ref, _ = load_golden("fixture3607")
    zipfile.ZipFile("/root/reference/tests/test_model_data.zip").extractall(tmp_path)
    cfg = yaml.safe_load(open(tmp_path / "config.yml"))
    keep = {k: v for k, v in cfg["nodes"].items()
            if v["type_"] in ("Land", "Sewer", "Groundwater", "River", "Node", "Waste") and "foul" not in k}
    cfg["nodes"] = keep
    cfg["arcs"] = {k: a for k, a in cfg["arcs"].items() if a["in_port"] in keep and a["out_port"] in keep}
    cfg["dates"] = cfg["dates"][:fm_ref.T]
    cfg.pop("extensions", None)
    yaml.safe_dump(cfg, open(tmp_path / "reduced.yml", "w"))
    m = Model()
    with contextlib.redirect_stdout(io.StringIO()):
        m.load(str(tmp_path), config_name="reduced.yml")
    fm = m.flatten()
    assert fm.node_names == fm_ref.node_names and fm.arc_names == fm_ref.arc_names
    for name in ("node_type", "node_gidx", "arc_in", "arc_out", "cand_ptr", "cand_arcs", "sched_op", "sched_node",
                 "month_idx", "dayofyear", "doy_adj", "land_surf_ptr", "land_surf"):
        assert np.array_equal(getattr(fm, name), getattr(fm_ref, name)), name
    for g in fm.groups:
        assert np.array_equal(fm.groups[g], fm_ref.groups[g]), g
    for name in ("node_forcing", "surf_inputs", "catch_inputs"):
        assert np.array_equal(getattr(fm, name), getattr(fm_ref, name)), name


@pytest.mark.reference
def test_oxford_demo_built_with_this_package_flattens_like_the_reference_build():
    """``wsi_b200.demo.create_oxford.create_oxford_model`` on the reference's demo data -> the very
    flat model tests/golden/oxford.npz was flattened from the REFERENCE's create_oxford_model
    (18 nodes / 21 arcs: FWTW, WWTW, ResidentialDemand, Reservoir abstraction, Sewer time-area)."""
    from wsi_b200.core import constants
    from wsi_b200.demo.create_oxford import create_oxford_model

    fm_ref, _ = load_golden("oxford")
    try:
        m = create_oxford_model("/root/reference/docs/demo/data")
        fm = m.flatten()
    finally:
        constants.set_default_pollutants()
        constants.FLOAT_ACCURACY = 1e-11
    assert fm.pollutants == fm_ref.pollutants and sorted(fm.additive) == sorted(fm_ref.additive)
    assert fm.node_names == fm_ref.node_names and fm.arc_names == fm_ref.arc_names
    for name in ("node_type", "node_gidx", "arc_in", "arc_out", "in_ptr", "in_arcs", "out_ptr", "cand_ptr", "cand_arcs",
                 "sched_op", "sched_node"):
        assert np.array_equal(getattr(fm, name), getattr(fm_ref, name)), name
    for g in fm.groups:
        assert np.array_equal(fm.groups[g], fm_ref.groups[g]), g
    for name in ("node_forcing", "catch_inputs"):
        assert np.array_equal(getattr(fm, name), getattr(fm_ref, name)), name


@pytest.mark.reference
def test_model_load_of_the_unreduced_reference_fixture(tmp_path):
    """The reference's real-catchment fixture AS IT IS (14 nodes: ResidentialDemand x2 on an
    UnlimitedDistribution, WWTW, storm + foul Sewers, garden irrigation, River, Groundwater ...)
    through this package's ``Model.load``: same flat model as the reference's own load gives."""
    import contextlib
    import io
    import zipfile

    import yaml

    from wsi_b200.orchestration.model import Model

    fm_ref, _ = load_golden("fixture3607_full")
    zipfile.ZipFile("/root/reference/tests/test_model_data.zip").extractall(tmp_path)
    cfg = yaml.safe_load(open(tmp_path / "config.yml"))
    cfg["dates"] = cfg["dates"][:fm_ref.T]
    cfg.pop("extensions", None)
    yaml.safe_dump(cfg, open(tmp_path / "full.yml", "w"))
    m = Model()
    with contextlib.redirect_stdout(io.StringIO()):
        m.load(str(tmp_path), config_name="full.yml")
    fm = m.flatten()
    assert fm.node_names == fm_ref.node_names and fm.arc_names == fm_ref.arc_names
    for name in ("node_type", "node_gidx", "arc_in", "arc_out", "in_ptr", "in_arcs", "cand_ptr", "cand_arcs", "sched_op",
                 "sched_node", "doy_adj"):
        assert np.array_equal(getattr(fm, name), getattr(fm_ref, name)), name
    for g in fm.groups:
        assert np.array_equal(fm.groups[g], fm_ref.groups[g]), g
    for name in ("node_forcing", "surf_inputs"):
        assert np.array_equal(getattr(fm, name), getattr(fm_ref, name)), name


def test_pickle_round_trip_and_result_files(tmp_path):
    """save_pickle / load_pickle (model.py:604-640) keep parameters and state; write_results writes
    the three csv files of the reference's command line (__main__.py:59-63)."""
    import pandas as pd

    from wsi_b200.orchestration.model import Model

    m, _ = _quickstart_model(T=5)
    m.nodes["my_groundwater"].tank.storage["volume"] = 12.5
    m.save_pickle(tmp_path / "m.pkl")
    m2 = Model().load_pickle(tmp_path / "m.pkl")
    assert list(m2.nodes) == list(m.nodes) and list(m2.arcs) == list(m.arcs)
    assert m2.nodes["my_groundwater"].tank.storage["volume"] == 12.5
    assert np.array_equal(m2.flatten().groups["tank"], m.flatten().groups["tank"])
    flows = [{"arc": "runoff", "flow": 0.05, "time": "2009-03-03", "phosphate": 1e-7}]
    Model.write_results(flows, [{"node": "my_sewer", "storage": 0.03, "time": "2009-03-03"}], [], tmp_path / "out")
    got = pd.read_csv(tmp_path / "out" / "flows.csv", index_col=0)
    assert got.loc[0, "arc"] == "runoff" and got.loc[0, "flow"] == 0.05
    assert (tmp_path / "out" / "tanks.csv").exists() and (tmp_path / "out" / "surfaces.csv").exists()


def test_pollutant_sets_without_kernels_are_padded_with_zero_additive_planes():
    """libwsi_b200.so has kernels for 2, 12 and 15 pollutants (layout.COMPILED_P); flatten pads any other
    count with additive pollutants that are zero everywhere, and neither results nor the objects'
    scattered-back state ever show them (goldens quickstart_p5 / quickstart_p14 pin the numbers)."""
    from oracle import oracle
    from wsi_b200.backend import flatten as F
    from wsi_b200.backend import layout as L
    from wsi_b200.backend.probes import ProbeSet
    from wsi_b200.core import constants
    from wsi_b200.orchestration.model import _RecordPlan

    m, _ = _quickstart_model(T=20)
    try:
        constants.POLLUTANTS = ["phosphate", "solids", "bod", "temperature", "ph"]
        constants.NON_ADDITIVE_POLLUTANTS = ["temperature", "ph"]
        constants.ADDITIVE_POLLUTANTS = ["phosphate", "solids", "bod"]
        m2, _ = None, None
        from wsi_b200.orchestration.model import Model

        m2 = Model()
        m2.add_nodes([
            {"type_": "Land", "data_input_dict": m.nodes["my_land"].data_input_dict, "name": "my_land", "surfaces": [
                {"type_": "ImperviousSurface", "surface": "urban", "area": 10, "pollutant_load": {"solids": 1e-7}}]},
            {"type_": "Sewer", "capacity": 0.04, "name": "my_sewer"},
            {"type_": "Waste", "name": "my_outlet"}])
        m2.add_arcs([{"type_": "Arc", "name": "drain", "in_port": "my_land", "out_port": "my_sewer"},
                     {"type_": "Arc", "name": "out", "in_port": "my_sewer", "out_port": "my_outlet"}])
        m2.dates = m.dates
        fm = m2.flatten()
        assert fm.P == 12 and fm.P in L.COMPILED_P
        assert F.model_pollutants(fm) == constants.POLLUTANTS
        assert set(fm.pollutants[5:]) <= set(fm.additive) and all(p.startswith(F.PAD_PREFIX) for p in fm.pollutants[5:])
        ps = ProbeSet(fm)
        ps.add_flows(list(m2.arcs))
        run = oracle.run(fm, probes=ps.probes)
        flows, _, _ = _RecordPlan(m2, fm, ps, [], [], False).records(ps.split(run.out_probe), m2.dates)
        assert set(flows[0]) == {"arc", "flow", "time", *constants.POLLUTANTS}
        assert not run.out_probe[:, :, 0].reshape(20, -1, fm.V)[:, :, 6:].any()  # padding planes stay zero
        F.scatter_back(m2, fm, run.groups, m2.dates[-1])
        assert set(m2.nodes["my_sewer"].sewer_tank.storage) == {"volume", *constants.POLLUTANTS}
        with pytest.raises(F.UnsupportedModelError):
            F.flatten(m2, m2.dates, [f"p{i}" for i in range(16)], [f"p{i}" for i in range(16)], [], 1e-11)
    finally:
        constants.set_default_pollutants()
        constants.FLOAT_ACCURACY = 1e-11


def test_arc_variants_and_river_reservoir_through_the_mirror_api():
    """PullArc / PushArc / SewerArc / WeirArc (arcs.py:808-905) and RiverReservoir (storage.py:992-1133) are
    accepted by Model.add_nodes / add_arcs and flatten to what golden 'reservoirs' (built with the
    reference's classes) holds; the queueing arcs are still refused."""
    import pandas as pd

    from wsi_b200.backend import layout as L
    from wsi_b200.backend.flatten import UnsupportedModelError
    from wsi_b200.core import constants
    from wsi_b200.orchestration.model import Model

    constants.set_default_pollutants()
    constants.FLOAT_ACCURACY = 1e-11
    fm_ref, _ = load_golden("reservoirs")
    dates = [pd.Timestamp("2003-01-01") + pd.Timedelta(days=i) for i in range(5)]
    tdata = {("temperature", d): 10.0 for d in dates}
    m = Model()
    m.add_nodes([
        {"type_": "River", "name": "r1", "length": 6000, "width": 10, "data_input_dict": tdata},
        {"type_": "Node", "name": "j"},
        {"type_": "RiverReservoir", "name": "rr", "capacity": 2.5e5, "area": 4e4, "environmental_flow": 9e3},
        {"type_": "Reservoir", "name": "farm", "capacity": 6e4, "area": 1e4},
        {"type_": "Reservoir", "name": "pond", "capacity": 3e4, "area": 1e4},
        {"type_": "Waste", "name": "out"}])
    m.add_arcs([
        {"type_": "Arc", "name": "r1-j", "in_port": "r1", "out_port": "j"},
        {"type_": "Arc", "name": "j-rr", "in_port": "j", "out_port": "rr"},
        {"type_": "PullArc", "name": "j-farm", "in_port": "j", "out_port": "farm"},
        {"type_": "PushArc", "name": "j-pond", "in_port": "j", "out_port": "pond"},
        {"type_": "WeirArc", "name": "rr-out", "in_port": "rr", "out_port": "out"}])
    m.orchestration.append({"RROnly": "satisfy_environmental"})
    m.nodes_type["RROnly"] = {"rr": m.nodes["rr"]}
    m.dates = dates
    fm = m.flatten()
    assert list(fm.arc_flags) == [0, 0, L.ARC_DENY_PUSH, L.ARC_DENY_PULL, 0]
    assert fm.node_type[2] == L.NT["RiverReservoir"] == fm_ref.node_type[4]
    assert fm.view("tank", "environmental_flow")[0, int(fm.node_gidx[2]), 0] == 9e3
    assert L.OP["RESERVOIR_ENVIRONMENTAL"] in list(fm.sched_op)
    # j pushes to [Node, River, Waste, Reservoir]: the PullArc is not a candidate
    nrwr = L.FI["NRWR"]
    cand = list(fm.cand_arcs[fm.cand_ptr[1 * L.NF + nrwr]: fm.cand_ptr[1 * L.NF + nrwr + 1]])
    assert cand == [1, 3]
    with pytest.raises(UnsupportedModelError):
        m.add_arcs([{"type_": "QueueArc", "name": "q", "in_port": "r1", "out_port": "j"}])


def test_distribution_and_base_demand_through_the_mirror_api():
    """Distribution with leakage (distribution.py:14-141) and the base Demand's constant demand
    (demand.py:14-131), built with this package's classes, flatten to what golden 'distribution' -- built
    with the reference's classes -- holds."""
    import pandas as pd

    from wsi_b200.backend import layout as L
    from wsi_b200.core import constants
    from wsi_b200.orchestration.model import Model

    constants.set_default_pollutants()
    constants.FLOAT_ACCURACY = 1e-11
    fm_ref, _ = load_golden("distribution")
    T = fm_ref.T
    dates = [pd.Timestamp("2005-01-01") + pd.Timedelta(days=i) for i in range(T)]
    tdata = {("temperature", d): float(fm_ref.node_forcing[t, 0, 2]) for t, d in enumerate(dates)}
    init = lambda v: {"volume": v, "do": 0.008, "org-phosphorus": v * 2e-5, "phosphate": v * 5e-5,  # noqa: E731
                      "ammonia": v * 1e-4, "solids": v * 1e-3, "bod": v * 1e-3, "cod": v * 2e-3, "ph": 7.1,
                      "temperature": 9.0, "nitrate": v * 2e-3, "nitrite": v * 1e-5, "org-nitrogen": v * 3e-4}
    house = {"do": 0.004, "org-phosphorus": 2e-7, "phosphate": 5e-7, "ammonia": 3e-6, "solids": 1e-5, "bod": 2e-5,
             "cod": 4e-5, "ph": 7.4, "temperature": 14.0, "nitrate": 1e-6, "nitrite": 1e-8, "org-nitrogen": 2e-6}
    m = Model()
    m.add_nodes([
        {"type_": "Reservoir", "name": "res", "capacity": 5e6, "area": 1e5, "initial_storage": init(7e5)},
        {"type_": "FWTW", "name": "fwtw", "service_reservoir_storage_capacity": 6e3, "service_reservoir_storage_area": 400,
         "service_reservoir_initial_storage": init(3e3), "treatment_throughput_capacity": 3.2e3, "data_input_dict": tdata},
        {"type_": "Distribution", "name": "mains", "leakage": 0.08},
        {"type_": "Groundwater", "name": "gw", "area": 1e5, "capacity": 4e3, "residence_time": 12.0, "initial_storage": init(3.2e3)},
        {"type_": "Demand", "name": "works", "constant_demand": 650.0,
         "pollutant_load": {"solids": 0.4, "bod": 0.6, "temperature": 21.0, "ph": 7.9}},
        {"type_": "Demand", "node_type_override": "ResidentialDemand", "name": "houses", "population": 1.6e4, "per_capita": 0.14,
         "pollutant_load": house, "data_input_dict": tdata},
        {"type_": "Sewer", "name": "foul", "capacity": 2.6e3, "pipe_timearea": {0: 0.7, 1: 0.3}},
        {"type_": "Node", "name": "j"},
        {"type_": "Waste", "name": "out"}])
    arcs = [("res-fwtw", "res", "fwtw", {}), ("fwtw-mains", "fwtw", "mains", {}), ("fwtw-foul", "fwtw", "foul", {}),
            ("mains-works", "mains", "works", {"capacity": 640.0}), ("mains-houses", "mains", "houses", {}),
            ("mains-gw", "mains", "gw", {}), ("works-foul", "works", "foul", {}), ("houses-foul", "houses", "foul", {}),
            ("foul-j", "foul", "j", {}), ("gw-j", "gw", "j", {}), ("j-out", "j", "out", {})]
    m.add_arcs([dict({"type_": "Arc", "name": a, "in_port": i, "out_port": o}, **kw) for a, i, o, kw in arcs])
    m.dates = dates
    fm = m.flatten()
    assert fm.node_type[2] == L.NT["Distribution"] and fm.view("distrib", "leakage")[0, 0, 0] == 0.08
    assert fm.view("demand", "population")[0, 0, 0] == -1 and fm.view("demand", "per_capita")[0, 0, 0] == 650.0
    for g in fm.groups:
        assert np.array_equal(fm.groups[g], fm_ref.groups[g]), g
    for k in ("node_type", "node_gidx", "arc_in", "arc_out", "cand_ptr", "cand_arcs", "sched_op", "sched_node", "arc_flags"):
        assert np.array_equal(getattr(fm, k), getattr(fm_ref, k)), k


@pytest.mark.reference
def test_patched_reference_objects_are_refused():
    """SURVEY 8b eligibility: reference objects whose behaviour was changed after construction -- by
    ``wsimod.extensions.apply_patches`` (extensions.py:94-119), by a user function assigned as a method or put
    into a handler dict or a surface's function lists, or by re-pointing a handler at another reference method --
    are refused by the flattener; the pristine model flattens (subprocess: the reference is imported)."""
    import subprocess
    import sys

    code = r"""
import sys; sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests/golden'); sys.dont_write_bytecode = True
import make_golden as G
from wsimod import extensions
from wsi_b200 import ensemble as E
from wsi_b200.backend.flatten import UnsupportedModelError
dates = E.daily_dates(20)
forcing = E.synthetic_forcing_arrays(dates, 1)
G.set_default_pollutants()
fdict = {(v, d): float(forcing[t, 0, j]) for t, d in enumerate(dates) for j, v in enumerate(('precipitation', 'et0', 'temperature'))}
def fresh():
    extensions.extensions_registry.clear()
    return G.config2_member(fdict, dates, 1.0, 1.0)
def refused(m, what):
    try:
        G.flatten_ref(m, dates)
    except UnsupportedModelError as e:
        assert what in str(e), (what, str(e))
        return
    raise AssertionError('accepted: ' + what)
m = fresh()
G.flatten_ref(m, dates)
river = 'gw'  # the member's Groundwater
land = next(n for n in m.nodes.values() if n.__class__.__name__ == 'Land')
# 1. a method patch through the reference's own registry
@extensions.register_node_patch(river, 'distribute')
def slower(self):
    return None
extensions.apply_patches(m)
refused(m, 'was patched')
# 2. a handler replaced through the registry (item patch), no _patched_ marker left behind
m = fresh()
@extensions.register_node_patch(river, 'push_set_handler', item='default')
def swallow(self, vqip):
    return self.empty_vqip()
extensions.apply_patches(m)
refused(m, 'user function')
# 3. a user function assigned as a method by hand
m = fresh()
m.nodes[river].infiltrate = lambda: None
refused(m, 'user function')
# 4. a handler re-pointed at another method of the reference
m = fresh()
m.nodes[river].push_set_handler['default'] = m.nodes[river].push_set_deny
refused(m, 'differ from a pristine Groundwater')
# 5. a function appended to a surface's list; 6. a reference method that does not belong to the class's lists
m = fresh()
s = m.nodes[land.name].surfaces[0]
s.processes.append(lambda: (s.empty_vqip(), s.empty_vqip()))
refused(m, 'user function')
m = fresh()
s = m.nodes[land.name].surfaces[0]
s.inflows.append(s.get_avail)
refused(m, 'function lists hold')
# 7. a tank whose end_timestep was replaced
m = fresh()
m.nodes[river].tank.end_timestep = lambda: None
refused(m, 'user function')
print('refused all')
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "refused all" in r.stdout, r.stdout + r.stderr


@pytest.mark.reference
def test_handler_table_is_current():
    """backend/handler_table.py (what a pristine reference object carries) is what tools/gen_handler_table.py
    derives from the mounted reference."""
    import subprocess
    import sys

    code = ("import sys; sys.path.insert(0, '/root/repo/tools'); sys.dont_write_bytecode = True\n"
            "import gen_handler_table as G\n"
            "from wsi_b200.backend.handler_table import HANDLER_TABLE\n"
            "assert G.build_table() == HANDLER_TABLE\nprint('current')\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "current" in r.stdout, r.stdout + r.stderr
