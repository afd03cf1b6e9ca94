#!/usr/bin/env python
"""Generate golden input/output fixtures by RUNNING THE REFERENCE (WSIMOD).

Only runs in the build container, where the reference is mounted read-only at
/root/reference.  For each scenario it

  1. builds the model with the reference's own classes,
  2. flattens the *initial* state with wsi_b200's flattener (duck-typed),
  3. replays ``Model.run``'s loop (wsimod/orchestration/model.py:993-1176) through
     the reference's public methods, capturing per step every arc VQIP, every
     node's (in, ds, out) mass-balance terms and all tank / surface / pool state,
  4. writes ``tests/golden/<name>.npz`` (flat model + reference outputs).

The GPU box has no /root/reference: tests read only the committed .npz files.

    PYTHONHASHSEED=0 PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py [names...]
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import tempfile
import zipfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.dont_write_bytecode = True

import pandas as pd  # noqa: E402
import yaml  # noqa: E402
from wsimod.core import constants  # noqa: E402
from wsimod.orchestration.model import Model, to_datetime  # noqa: E402

from wsi_b200.backend import flatten as F  # noqa: E402
from wsi_b200.backend import layout as L  # noqa: E402
from wsi_b200.backend.probes import SURF_OUT, SURF_POOLS, SURF_STATE  # noqa: E402


# ------------------------------------------------------------------ harness
def set_default_pollutants():
    constants.set_default_pollutants()
    constants.FLOAT_ACCURACY = 1e-11


def vq(d):
    return [d["volume"]] + [d[p] for p in constants.POLLUTANTS]


def run_reference(model, dates):
    """Replay Model.run (model.py:993-1176) and capture everything per step."""
    nodes = list(model.nodes.values())
    arcs = list(model.arcs.values())
    lands = [n for n in nodes if n.__class__.__name__ == "Land"]
    surfs = [s for n in lands for s in n.surfaces]
    stores = [n for n in nodes if n.__class__.__name__ in F.STORE_TYPES]
    sewers = [n for n in nodes if n.__class__.__name__ == "Sewer"]
    V = len(constants.POLLUTANTS) + 1
    T = len(dates)
    out = {
        "flows": np.zeros((T, len(arcs), V)),
        "mb": np.zeros((T, len(nodes), 3, V)),
        "land_tanks": np.zeros((T, len(lands), 3 * V)),
        "surf_storage": np.zeros((T, len(surfs), V)),
        "surf_out": np.zeros((T, len(surfs), len(SURF_OUT))),
        "surf_state": np.zeros((T, len(surfs), len(SURF_STATE) + 10)),
        "tank_storage": np.zeros((T, len(stores), V)),
        "sewer_storage": np.zeros((T, len(sewers), V)),
        "sewer_active": np.zeros((T, len(sewers), V)),
    }
    advisories = io.StringIO()
    with contextlib.redirect_stdout(advisories):
        for t, date in enumerate(dates):
            for node in model.nodelist:
                node.t = date
                node.monthyear = date.to_period("M")
            for item in model.orchestration:
                for node_type, function in item.items():
                    for node in model.nodes_type.get(node_type, {}).values():
                        getattr(node, function)()
            for name in model.river_discharge_order:
                model.nodes[name].distribute()
            for arc in arcs:
                arc.arc_mass_balance()
            for i, node in enumerate(nodes):
                in_, ds_, out_ = node.node_mass_balance()
                out["mb"][t, i, 0] = vq(in_)
                out["mb"][t, i, 1] = vq(ds_)
                out["mb"][t, i, 2] = vq(out_)
            for i, arc in enumerate(arcs):
                out["flows"][t, i] = vq(arc.vqip_out)
            for i, n in enumerate(lands):
                out["land_tanks"][t, i] = (
                    vq(n.surface_runoff.storage) + vq(n.subsurface_runoff.storage) + vq(n.percolation.storage)
                )
            for i, s in enumerate(surfs):
                out["surf_storage"][t, i] = vq(s.storage)
                g = lambda a: getattr(s, a)["volume"] if hasattr(s, a) else 0.0  # noqa: E731
                out["surf_out"][t, i] = [
                    g("percolation"), g("subsurface_flow"), g("infiltration_excess"), g("evaporation"),
                    g("precipitation"), getattr(s, "tank_recharge", 0.0),
                ]
                st = [
                    getattr(s, "et0_coefficient", 0.0),
                    (-1.0 if s.days_after_sow is None else s.days_after_sow) if hasattr(s, "days_after_sow") else 0.0,
                    getattr(s, "crop_cover", 0.0), getattr(s, "ground_cover", 0.0), getattr(s, "crop_factor", 0.0),
                ]
                if hasattr(s, "nutrient_pool"):
                    npool = s.nutrient_pool
                    for pool in (npool.fast_pool, npool.humus_pool, npool.dissolved_inorganic_pool,
                                 npool.dissolved_organic_pool, npool.adsorbed_inorganic_pool):
                        st += [pool.storage["N"], pool.storage["P"]]
                else:
                    st += [0.0] * 10
                out["surf_state"][t, i] = st
            for i, n in enumerate(stores):
                out["tank_storage"][t, i] = vq(n.tank.storage)
            for i, n in enumerate(sewers):
                out["sewer_storage"][t, i] = vq(n.sewer_tank.storage)
                out["sewer_active"][t, i] = vq(n.sewer_tank.active_storage)
            for node in nodes:
                node.end_timestep()
            for arc in arcs:
                arc.end_timestep()
    out["n_advisories"] = np.array([len(advisories.getvalue().splitlines())])
    return out


def flatten_ref(model, dates):
    return F.flatten(
        model, dates, constants.POLLUTANTS, constants.ADDITIVE_POLLUTANTS,
        constants.NON_ADDITIVE_POLLUTANTS, constants.FLOAT_ACCURACY,
    )


def merge_replicas(fms):
    """Stack R=1 flat models of identical structure into one R=len(fms) model."""
    base = fms[0]
    groups = {}
    for g in base.groups:
        cols = [fm.groups[g] for fm in fms]  # each [nf, n_obj]
        groups[g] = np.ascontiguousarray(np.stack(cols, axis=2).reshape(cols[0].shape[0], -1))
    kw = dict(base.__dict__)
    kw.update(groups=groups, R=len(fms))
    return F.FlatModel(**kw)


def pad_ref(ref, V_ref, V):
    """The flat model pads its pollutant set up to a compiled width (flatten.py): the reference's
    outputs get the same -- all-zero -- planes, so a comparison also checks that they stay zero."""
    if V == V_ref:
        return ref
    out = dict(ref)
    grow = lambda a: np.concatenate([a, np.zeros(a.shape[:-1] + (V - V_ref,))], axis=-1)  # noqa: E731
    for k in ("flows", "mb", "surf_storage", "tank_storage", "sewer_storage", "sewer_active"):
        out[k] = grow(ref[k])
    lt = ref["land_tanks"]
    out["land_tanks"] = grow(lt.reshape(lt.shape[:-1] + (3, V_ref))).reshape(lt.shape[:-1] + (3 * V,))
    return out


def save(name, fm, ref, note):
    d = fm.to_npz_dict()
    ref = pad_ref(ref, ref["flows"].shape[-1], fm.V)
    for k, v in ref.items():
        d["ref_" + k] = v
    d["meta_note"] = np.array(note)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **d)
    print(f"wrote {path}: {os.path.getsize(path) / 1e6:.2f} MB")


# ------------------------------------------------------------------ scenarios
def oxford_land_inputs():
    df = pd.read_csv(os.path.join(REF, "docs/demo/data/processed/timeseries_data.csv"))
    df = df.loc[df.site == "oxford_land"].copy()
    df.loc[df.variable == "precipitation", "value"] *= constants.MM_TO_M
    df.date = pd.to_datetime(df.date)
    dates = sorted(df.date.drop_duplicates())
    return df.set_index(["variable", "date"]).value.to_dict(), [pd.Timestamp(x) for x in dates]


def build_quickstart(land_inputs, load_pol="phosphate"):
    """docs/demo/scripts/quickstart_demo.py:84-190."""
    sewer = {"type_": "Sewer", "capacity": 0.04, "name": "my_sewer"}
    surface1 = {"type_": "ImperviousSurface", "surface": "urban", "area": 10, "pollutant_load": {load_pol: 1e-7}}
    surface2 = {"type_": "PerviousSurface", "surface": "rural", "area": 100, "depth": 0.5,
                "pollutant_load": {load_pol: 1e-7}}
    land = {"type_": "Land", "data_input_dict": land_inputs, "surfaces": [surface1, surface2], "name": "my_land"}
    gw = {"type_": "Groundwater", "area": 100, "capacity": 100, "name": "my_groundwater"}
    node = {"type_": "Node", "name": "my_river"}
    waste = {"type_": "Waste", "name": "my_outlet"}
    arcs = [
        ("urban_drainage", "my_land", "my_sewer"),
        ("percolation", "my_land", "my_groundwater"),
        ("runoff", "my_land", "my_river"),
        ("storm_outflow", "my_sewer", "my_river"),
        ("baseflow", "my_groundwater", "my_river"),
        ("catchment_outflow", "my_river", "my_outlet"),
    ]
    m = Model()
    m.add_nodes([sewer, land, gw, node, waste])
    m.add_arcs([{"type_": "Arc", "name": a, "in_port": i, "out_port": o} for a, i, o in arcs])
    return m


def scenario_quickstart(T=1456):
    set_default_pollutants()
    land_inputs, dates = oxford_land_inputs()
    dates = dates[:T]
    m = build_quickstart(land_inputs)
    fm = flatten_ref(m, dates)
    ref = run_reference(m, dates)
    save("quickstart", fm, ref, "quickstart_demo.py model, default pollutants, first %d days" % T)


def scenario_quickstart_simple(T=300):
    constants.set_simple_pollutants()
    constants.FLOAT_ACCURACY = 1e-11
    land_inputs, dates = oxford_land_inputs()
    dates = dates[200 : 200 + T]
    m = build_quickstart(land_inputs)
    fm = flatten_ref(m, dates)
    ref = run_reference(m, dates)
    save("quickstart_simple", fm, ref, "quickstart model with set_simple_pollutants (P=2)")
    set_default_pollutants()


def scenario_quickstart_p5(T=400):
    """Five pollutants (three additive, two not): a count the library has no kernels for, so the
    flat model is padded to twelve (flatten.py) -- the reference runs the five."""
    constants.POLLUTANTS = ["phosphate", "solids", "bod", "temperature", "ph"]
    constants.NON_ADDITIVE_POLLUTANTS = ["temperature", "ph"]
    constants.ADDITIVE_POLLUTANTS = ["phosphate", "solids", "bod"]
    constants.FLOAT_ACCURACY = 1e-11
    land_inputs, dates = oxford_land_inputs()
    dates = dates[100 : 100 + T]
    m = build_quickstart(land_inputs, load_pol="solids")
    fm = flatten_ref(m, dates)
    ref = run_reference(m, dates)
    save("quickstart_p5", fm, ref, "quickstart model with 5 pollutants (padded to 12 in the flat model)")
    set_default_pollutants()


def scenario_quickstart_p14(T=300):
    """The default twelve pollutants plus two more additive ones: padded to fifteen in the flat model."""
    constants.set_default_pollutants()
    constants.POLLUTANTS = list(constants.POLLUTANTS) + ["chloride", "zinc"]
    constants.ADDITIVE_POLLUTANTS = list(constants.ADDITIVE_POLLUTANTS) + ["chloride", "zinc"]
    constants.FLOAT_ACCURACY = 1e-11
    land_inputs, dates = oxford_land_inputs()
    dates = dates[300 : 300 + T]
    m = build_quickstart(land_inputs, load_pol="zinc")
    fm = flatten_ref(m, dates)
    ref = run_reference(m, dates)
    save("quickstart_p14", fm, ref, "quickstart model with 14 pollutants (padded to 15 in the flat model)")
    set_default_pollutants()


def scenario_fixture3607(T=730):
    """tests/test_model_data.zip reduced to the nodes on the GPU path (Land with 9
    surfaces incl. 7 GrowingSurface + GardenSurface, storm Sewer, Groundwater, River
    with biochemistry, junction, 2 Waste); csv date regime (to_datetime)."""
    set_default_pollutants()
    tmp = tempfile.mkdtemp()
    zipfile.ZipFile(os.path.join(REF, "tests/test_model_data.zip")).extractall(tmp)
    cfg = yaml.safe_load(open(os.path.join(tmp, "config.yml")))
    keep = {k: v for k, v in cfg["nodes"].items()
            if v["type_"] in ("Land", "Sewer", "Groundwater", "River", "Node", "Waste") and "foul" not in k}
    cfg["nodes"] = keep
    cfg["arcs"] = {k: a for k, a in cfg["arcs"].items() if a["in_port"] in keep and a["out_port"] in keep}
    cfg["dates"] = cfg["dates"][:T]
    cfg.pop("extensions", None)
    yaml.safe_dump(cfg, open(os.path.join(tmp, "reduced.yml"), "w"))
    m = Model()
    with contextlib.redirect_stdout(io.StringIO()):
        m.load(tmp, config_name="reduced.yml")
    dates = m.dates
    fm = flatten_ref(m, dates)
    ref = run_reference(m, dates)
    save("fixture3607", fm, ref, "reference test fixture 3607 reduced to supported nodes, %d days, csv dates" % T)


def scenario_hourly(T=1080):
    """Hourly timestep (config 5's regime): dates are the reference's own ``to_datetime`` objects
    parsed from 'YYYY-MM-DD HH:MM:SS' (model.py:139-162), 45 days from 2001-03-06.  The crop calendar
    advances per TIMESTEP (land.py:1531-1544): days_after_sow counts hours, sowing and harvest fire on
    the first hour of their day, and -- to_datetime.is_leap_year being a method -- the day of year is
    shifted in this non-leap year too."""
    from datetime import datetime, timedelta

    set_default_pollutants()
    t0 = datetime(2001, 3, 6)
    dates = [to_datetime((t0 + timedelta(hours=i)).strftime("%Y-%m-%d %H:%M:%S")) for i in range(T)]
    rng = np.random.default_rng(20261105)
    hours = np.arange(T)
    temperature = 8 + 5 * np.sin(2 * np.pi * (hours % 24 - 9) / 24) + rng.normal(0, 0.7, T)
    et0 = np.maximum(0.0, 2.5e-4 * np.sin(2 * np.pi * (hours % 24 - 6) / 24))
    wet = rng.random(T) < 0.12
    precipitation = wet * rng.gamma(0.7, 1.2e-3, T)
    forcing = {}
    for t, d in enumerate(dates):
        forcing[("precipitation", d)] = float(precipitation[t])
        forcing[("et0", d)] = float(et0[t])
        forcing[("temperature", d)] = float(temperature[t])
    monthly = {(var, to_datetime(m)): 1e-7 / 24 for var in L.DEP_COLS for m in ("2001-03", "2001-04")}
    a_grow = 3.0e4
    vol = a_grow * 1.25 * 0.4 * 0.9
    init = {"do": 0.01, "org-phosphorus": vol * 1e-4, "phosphate": vol * 1e-4, "ammonia": vol * 5e-5,
            "solids": vol * 1e-2, "bod": vol * 1e-3, "cod": vol * 2e-2, "ph": 7.0, "temperature": 9.0,
            "nitrate": vol * 4e-3, "nitrite": vol * 2e-5, "org-nitrogen": vol * 5e-3, "volume": vol}
    soil = {"org-phosphorus": vol * 5e-3, "phosphate": vol * 5e-3, "ammonia": 0.0, "nitrate": vol,
            "nitrite": 0.0, "org-nitrogen": vol, "volume": vol}
    grow = dict(GROWING_TEMPLATE)
    grow.update(type_="GrowingSurface", surface="crop", area=a_grow, pollutant_load={}, initial_storage=init,
                initial_soil_storage=soil, data_input_dict=monthly, sowing_day=74, harvest_day=95,
                crop_factor_stage_dates=[0, 73, 74, 80, 86, 90, 95, 96, 366], infiltration_capacity=0.004,
                decays={"ammonia": {"constant": 0.1 / 24, "exponent": 1.005},
                        "nitrite": {"constant": 0.05 / 24, "exponent": 1.005}})
    imp = {"type_": "ImperviousSurface", "surface": "urban", "area": 5e3, "pore_depth": 0.002,
           "pollutant_load": {"phosphate": 1e-11, "solids": 1e-6}}
    land = {"type_": "Land", "name": "land", "data_input_dict": forcing, "surfaces": [imp, grow],
            "surface_residence_time": 6.0, "subsurface_residence_time": 60.0, "percolation_residence_time": 400.0}
    sewer = {"type_": "Sewer", "name": "sewer", "capacity": 40.0, "pipe_time": 2}
    gw = {"type_": "Groundwater", "name": "gw", "area": a_grow, "capacity": a_grow * 0.5, "residence_time": 900.0,
          "initial_storage": a_grow * 0.01}
    arcs = [("land-sewer", "land", "sewer"), ("land-gw", "land", "gw"), ("land-river", "land", "river"),
            ("sewer-river", "sewer", "river"), ("gw-river", "gw", "river"), ("river-outlet", "river", "outlet")]
    m = Model()
    m.add_nodes([sewer, land, gw, {"type_": "Node", "name": "river"}, {"type_": "Waste", "name": "outlet"}])
    m.add_arcs([{"type_": "Arc", "name": a, "in_port": i, "out_port": o} for a, i, o in arcs])
    fm = flatten_ref(m, dates)
    ref = run_reference(m, dates)
    das = ref["surf_state"][:, 1, 1]
    assert das.max() > 24 * 15, "days_after_sow must count hours"
    save("hourly", fm, ref, "hourly timestep, to_datetime dates, crop calendar counting timesteps, %d steps" % T)


def scenario_fixture3607_full(T=730):
    """tests/test_model_data.zip as it is: 14 nodes / 21 arcs -- Land with 9 surfaces (GardenSurface
    irrigated by a ResidentialDemand through the ("Demand", "Garden") tag), storm and two foul Sewers
    exchanging through tiny-capacity arcs, Groundwater, River with biochemistry, WWTW, two
    ResidentialDemands drawing on an UnlimitedDistribution, junction, 3 Waste; csv date regime."""
    set_default_pollutants()
    tmp = tempfile.mkdtemp()
    zipfile.ZipFile(os.path.join(REF, "tests/test_model_data.zip")).extractall(tmp)
    cfg = yaml.safe_load(open(os.path.join(tmp, "config.yml")))
    cfg["dates"] = cfg["dates"][:T]
    cfg.pop("extensions", None)
    yaml.safe_dump(cfg, open(os.path.join(tmp, "full.yml"), "w"))
    m = Model()
    with contextlib.redirect_stdout(io.StringIO()):
        m.load(tmp, config_name="full.yml")
    dates = m.dates
    fm = flatten_ref(m, dates)
    ref = run_reference(m, dates)
    save("fixture3607_full", fm, ref, "reference test fixture 3607, every node, %d days, csv dates" % T)


def scenario_city(D=3, T=480):
    """BASELINE.json configs[4] at a size the reference runs in seconds: a synthetic city of D
    districts x {Land(Impervious, Garden, Pervious), storm Sewer, foul Sewer, Groundwater,
    ResidentialDemand (house -> foul sewer, garden -> Land)} on a shared chain of three Rivers, a
    Reservoir abstracting from the middle river (minimum required flow, capacity-limited arc), FWTW
    -> distribution junction -> demands, foul sewers -> WWTW -> river; HOURLY timestep
    (to_datetime dates, 'YYYY-MM-DD HH:MM:SS'), 20 days."""
    from datetime import datetime, timedelta

    set_default_pollutants()
    t0 = datetime(2001, 6, 1)
    dates = [to_datetime((t0 + timedelta(hours=i)).strftime("%Y-%m-%d %H:%M:%S")) for i in range(T)]
    rng = np.random.default_rng(20261106)
    hours = np.arange(T)
    months = [to_datetime("2001-06")]

    def land_forcing(k):
        temperature = 15 + 6 * np.sin(2 * np.pi * (hours % 24 - 9) / 24) + rng.normal(0, 0.5, T) + 0.3 * k
        et0 = np.maximum(0.0, 3.0e-4 * np.sin(2 * np.pi * (hours % 24 - 6) / 24))
        precipitation = (rng.random(T) < 0.10) * rng.gamma(0.7, 1.5e-3, T)
        d = {}
        for t, date in enumerate(dates):
            d[("precipitation", date)] = float(precipitation[t])
            d[("et0", date)] = float(et0[t])
            d[("temperature", date)] = float(temperature[t])
        return d

    tdata = {("temperature", d): float(14 + 2 * np.sin(2 * np.pi * (h % 24 - 12) / 24)) for h, d in zip(hours, dates)}
    monthly = {(var, m): 1e-7 / 24 for var in L.DEP_COLS for m in months}
    init = lambda v: {"volume": v, "do": 0.008, "org-phosphorus": v * 2e-5, "phosphate": v * 5e-5,  # noqa: E731
                      "ammonia": v * 1e-4, "solids": v * 1e-3, "bod": v * 1e-3, "cod": v * 2e-3, "ph": 7.1,
                      "temperature": 13.0, "nitrate": v * 2e-3, "nitrite": v * 1e-5, "org-nitrogen": v * 3e-4}
    house = {"do": 0.004, "org-phosphorus": 2e-7, "phosphate": 5e-7, "ammonia": 3e-6, "solids": 1e-5, "bod": 2e-5,
             "cod": 4e-5, "ph": 7.4, "temperature": 14.0, "nitrate": 1e-6, "nitrite": 1e-8, "org-nitrogen": 2e-6}
    nodes, arcs = [], []
    arc = lambda n, a, b, **kw: arcs.append(dict({"type_": "Arc", "name": n, "in_port": a, "out_port": b}, **kw))  # noqa: E731
    for k in range(D):
        a_g = 2.0e4 * (1 + 0.5 * k)
        vol = a_g * 1.25 * 0.4 * 0.9
        garden = dict(GROWING_TEMPLATE)
        garden.update(type_="GardenSurface", surface="garden", area=a_g, pollutant_load={}, data_input_dict=monthly,
                      initial_storage=dict(init(vol), volume=vol),
                      initial_soil_storage={"org-phosphorus": vol * 5e-3, "phosphate": vol * 5e-3, "ammonia": 0.0,
                                            "nitrate": vol, "nitrite": 0.0, "org-nitrogen": vol, "volume": vol},
                      sowing_day=1, harvest_day=365, infiltration_capacity=0.004,
                      decays={"ammonia": {"constant": 0.1 / 24, "exponent": 1.005},
                              "nitrite": {"constant": 0.05 / 24, "exponent": 1.005}})
        imp = {"type_": "ImperviousSurface", "surface": "urban", "area": 3e4 * (1 + k), "pore_depth": 0.002,
               "pollutant_load": {"phosphate": 1e-11, "solids": 1e-6}}
        perv = {"type_": "PerviousSurface", "surface": "park", "area": 4e4, "depth": 0.5,
                "pollutant_load": {"phosphate": 1e-9}, "initial_storage": 4e4 * 0.5 * 0.4 * 0.5,
                "infiltration_capacity": 0.003}
        nodes += [
            {"type_": "Land", "name": f"land{k}", "data_input_dict": land_forcing(k), "surfaces": [imp, garden, perv],
             "surface_residence_time": 4.0, "subsurface_residence_time": 48.0, "percolation_residence_time": 300.0},
            {"type_": "Sewer", "name": f"storm{k}", "capacity": 60.0 * (1 + k), "pipe_time": k % 2},
            {"type_": "Sewer", "name": f"foul{k}", "capacity": 500.0, "pipe_timearea": {0: 0.6, 1: 0.3, 2: 0.1}},
            {"type_": "Groundwater", "name": f"gw{k}", "area": 1e5, "capacity": 5e4, "residence_time": 700.0,
             "initial_storage": 2e3},
            {"type_": "Demand", "node_type_override": "ResidentialDemand", "name": f"demand{k}", "population": 4e3 * (1 + k), "per_capita": 0.15 / 24,
             "pollutant_load": house, "data_input_dict": tdata, "gardening_efficiency": 0.4},
        ]
        arc(f"land{k}-storm{k}", f"land{k}", f"storm{k}")
        arc(f"land{k}-gw{k}", f"land{k}", f"gw{k}")
        arc(f"land{k}-river{k}", f"land{k}", f"river{k}")
        arc(f"storm{k}-river{k}", f"storm{k}", f"river{k}", capacity=25.0 * (1 + k))
        arc(f"gw{k}-river{k}", f"gw{k}", f"river{k}")
        arc(f"dist-demand{k}", "distribution", f"demand{k}")
        arc(f"demand{k}-foul{k}", f"demand{k}", f"foul{k}")
        arc(f"demand{k}-land{k}", f"demand{k}", f"land{k}", preference=0.3)
        arc(f"foul{k}-wwtw", f"foul{k}", "wwtw", preference=1e10)
        arc(f"foul{k}-river{k}", f"foul{k}", f"river{k}", preference=1e-10)
    for k in range(3):
        nodes.append({"type_": "River", "name": f"river{k}", "length": 3000.0 * (1 + k), "width": 8.0 + 4 * k,
                      "data_input_dict": tdata, "initial_storage": init(4e3 * (1 + k)), "mrf": 20.0 if k == 1 else 0})
    for k in range(3, D):  # further districts drain to the last river
        pass
    nodes += [
        {"type_": "Reservoir", "name": "reservoir", "capacity": 4e5, "initial_storage": init(3.5e5), "area": 5e4, "datum": 10},
        {"type_": "FWTW", "name": "fwtw", "service_reservoir_storage_capacity": 1e4, "service_reservoir_storage_area": 500,
         "service_reservoir_initial_storage": 1.5e3, "treatment_throughput_capacity": 400.0},
        {"type_": "WWTW", "name": "wwtw", "stormwater_storage_capacity": 300.0, "stormwater_storage_area": 100,
         "treatment_throughput_capacity": 300.0},
        {"type_": "Node", "name": "distribution"},
        {"type_": "Waste", "name": "sea"},
    ]
    arc("river0-river1", "river0", "river1")
    arc("river1-river2", "river1", "river2")
    arc("river2-sea", "river2", "sea")
    arc("river1-reservoir", "river1", "reservoir", capacity=150.0)
    arc("reservoir-fwtw", "reservoir", "fwtw")
    arc("fwtw-dist", "fwtw", "distribution")
    arc("fwtw-foul0", "fwtw", "foul0")
    arc("wwtw-river2", "wwtw", "river2")
    m = Model()
    m.add_nodes(nodes)
    m.add_arcs(arcs)
    fm = flatten_ref(m, dates)
    ref = run_reference(m, dates)
    save("city", fm, ref, "synthetic city, %d districts, hourly, %d steps" % (D, T))


def scenario_queuegw(T=400):
    """QueueGroundwater (storage.py:248-449) both ways: with ``decays`` -- a DecayQueueTank whose
    DecayArcAlt decays what enters and what waits (tanks.py:645-696, arcs.py:731-805; built-in delay of
    one timestep) -- and without (a plain QueueTank), each fed by a Land through a time-area diagram
    and releasing to a river junction."""
    set_default_pollutants()
    dates = [pd.Timestamp("2001-01-01") + pd.Timedelta(days=i) for i in range(T)]
    rng = np.random.default_rng(20261107)
    forcing = synthetic_forcing(dates, rng)
    tdata = {k: v for k, v in forcing.items() if k[0] == "temperature"}
    nodes, arcs = [], []
    for k, decays in enumerate([{"ammonia": {"constant": 0.04, "exponent": 1.02}, "bod": {"constant": 0.1, "exponent": 1.0}},
                                None]):
        perv = {"type_": "PerviousSurface", "surface": "rural", "area": 5e4 * (1 + k), "depth": 0.5,
                "pollutant_load": {"phosphate": 1e-7, "ammonia": 2e-7, "bod": 1e-6}, "initial_storage": 5e4 * (1 + k) * 0.5 * 0.4 * 0.5,
                "infiltration_capacity": 0.05}
        nodes += [
            {"type_": "Land", "name": f"land{k}", "data_input_dict": forcing, "surfaces": [perv],
             "percolation_residence_time": 5.0},
            {"type_": "QueueGroundwater", "name": f"qgw{k}", "area": 1e5, "capacity": 4e4, "initial_storage": 3e3,
             "timearea": {0: 0.5, 1: 0.3, 3: 0.2}, "decays": decays, "data_input_dict": tdata},
        ]
        arcs += [(f"land{k}-qgw{k}", f"land{k}", f"qgw{k}"), (f"land{k}-river", f"land{k}", "river"),
                 (f"qgw{k}-river", f"qgw{k}", "river")]
    nodes += [{"type_": "Node", "name": "river"}, {"type_": "Waste", "name": "outlet"}]
    arcs.append(("river-outlet", "river", "outlet"))
    m = Model()
    m.add_nodes(nodes)
    m.add_arcs([{"type_": "Arc", "name": a, "in_port": i, "out_port": o} for a, i, o in arcs])
    # Node.__init__ resets data_input_dict to None after QueueGroundwater stored it (storage.py:296-297,
    # nodes.py:58), so the temperature its queue decays with has to be attached afterwards
    m.nodes["qgw0"].data_input_dict = tdata
    fm = flatten_ref(m, dates)
    ref = run_reference(m, dates)
    save("queuegw", fm, ref, "QueueGroundwater with and without decays, %d days" % T)


def synthetic_forcing(dates, rng):
    """SURVEY 8d config 2 forcing generator."""
    doy = np.array([d.dayofyear for d in dates], dtype=float)
    temperature = 10 + 8 * np.sin(2 * np.pi * (doy - 110) / 365) + rng.normal(0, 2, len(dates))
    et0 = np.maximum(2e-4, 1e-3 * (1.5 + 1.2 * np.sin(2 * np.pi * (doy - 100) / 365)))
    wet = rng.random(len(dates)) < 0.45
    precipitation = wet * rng.gamma(0.8, 6e-3, len(dates))
    d = {}
    for t, date in enumerate(dates):
        d[("precipitation", date)] = float(precipitation[t])
        d[("et0", date)] = float(et0[t])
        d[("temperature", date)] = float(temperature[t])
    return d


GROWING_TEMPLATE = dict(  # "Spring wheat" of the 3607 fixture
    surface_coefficient=0.05, field_capacity=0.3577337861061096, wilting_point=0.14462512731552124,
    crop_factor_stages=[0.0, 0.0, 0.3, 0.3, 1.15, 1.15, 0.325, 0.0, 0.0],
    crop_factor_stage_dates=[0, 73, 74, 121, 166, 196, 274, 275, 366],
    ihacres_p=10.0, rooting_depth=1.25, percolation_coefficient=0.25, total_porosity=0.4,
    ET_depletion_factor=0.55, sowing_day=74, harvest_day=274, infiltration_capacity=0.06,
)


def config2_member(forcing, dates, scale, area_scale):
    """One {Land(Impervious, Pervious, Growing) -> Sewer, Groundwater, Node -> Waste}
    catchment of SURVEY 8d config 2; ``scale`` perturbs parameters."""
    months = sorted({(d.year, d.month) for d in dates})
    monthly = {}
    for var in L.DEP_COLS:
        for (y, mth) in months:
            monthly[(var, pd.Timestamp(year=y, month=mth, day=1).to_period("M"))] = 1e-7
    a_imp, a_perv, a_grow = 2.0e4 * area_scale, 5.0e4 * area_scale, 2.0e5 * area_scale
    vol = a_grow * 1.25 * 0.4 * 0.9
    init = {"do": 0.01, "org-phosphorus": vol * 1e-4, "phosphate": vol * 1e-4, "ammonia": vol * 5e-5,
            "solids": vol * 1e-2, "bod": vol * 1e-3, "cod": vol * 2e-2, "ph": 7.0, "temperature": 11.0,
            "nitrate": vol * 4e-3, "nitrite": vol * 2e-5, "org-nitrogen": vol * 5e-3, "volume": vol}
    soil = {"org-phosphorus": vol * 5e-3, "phosphate": vol * 5e-3, "ammonia": 0.0, "nitrate": vol,
            "nitrite": 0.0, "org-nitrogen": vol, "volume": vol}
    grow = dict(GROWING_TEMPLATE)
    grow.update(
        type_="GrowingSurface", surface="crop", area=a_grow, pollutant_load={}, initial_storage=init,
        initial_soil_storage=soil, data_input_dict=monthly,
        decays={"ammonia": {"constant": 0.1 * scale, "exponent": 1.005},
                "nitrite": {"constant": 0.05 * scale, "exponent": 1.005}},
        field_capacity=GROWING_TEMPLATE["field_capacity"] * min(scale, 1.1),
        infiltration_capacity=0.06 * scale, rooting_depth=1.25 * scale,
    )
    imp = {"type_": "ImperviousSurface", "surface": "urban", "area": a_imp, "pore_depth": 0.004 * scale,
           "pollutant_load": {"phosphate": 1e-10, "solids": 1e-5, "bod": 1e-5, "ammonia": 1e-9,
                              "temperature": 13.0, "do": 0.005, "ph": 7.5}}
    perv = {"type_": "PerviousSurface", "surface": "rural", "area": a_perv, "depth": 0.5 * scale,
            "pollutant_load": {"phosphate": 1e-7}, "initial_storage": a_perv * 0.5 * scale * 0.4 * 0.5,
            "infiltration_capacity": 0.05}
    land = {"type_": "Land", "name": "land", "data_input_dict": forcing, "surfaces": [imp, perv, grow],
            "surface_residence_time": 2.0 * scale, "subsurface_residence_time": 10.0,
            "percolation_residence_time": 30.0 / scale}
    sewer = {"type_": "Sewer", "name": "sewer", "capacity": a_imp * 0.01 * scale}
    gw = {"type_": "Groundwater", "name": "gw", "area": a_perv + a_grow, "capacity": (a_perv + a_grow) * 0.5,
          "residence_time": 100.0 * scale, "initial_storage": (a_perv + a_grow) * 0.01}
    node = {"type_": "Node", "name": "river"}
    waste = {"type_": "Waste", "name": "outlet"}
    arcs = [("land-sewer", "land", "sewer"), ("land-gw", "land", "gw"), ("land-river", "land", "river"),
            ("sewer-river", "sewer", "river"), ("gw-river", "gw", "river"), ("river-outlet", "river", "outlet")]
    m = Model()
    m.add_nodes([sewer, land, gw, node, waste])
    m.add_arcs([{"type_": "Arc", "name": a, "in_port": i, "out_port": o} for a, i, o in arcs])
    return m


def scenario_config2(T=730, R=4):
    set_default_pollutants()
    dates = [pd.Timestamp("2000-01-01") + pd.Timedelta(days=i) for i in range(T)]
    rng = np.random.default_rng(20261019)
    forcing = synthetic_forcing(dates, rng)
    scales = [1.0, 0.85, 1.2, 0.95][:R]
    areas = [1.0, 0.3, 4.0, 12.0][:R]
    fms, refs = [], []
    for s, a in zip(scales, areas):
        m = config2_member(forcing, dates, s, a)
        fms.append(flatten_ref(m, dates))
        refs.append(run_reference(m, dates))
    fm = merge_replicas(fms)
    ref = {k: np.stack([r[k] for r in refs], axis=-1) for k in refs[0] if k != "n_advisories"}
    ref["n_advisories"] = np.array([int(sum(r["n_advisories"][0] for r in refs))])
    save("config2", fm, ref, "SURVEY 8d config 2 catchment, %d perturbed members, %d days" % (R, T))


def scenario_rivernet(T=420):
    """Catchments -> river network with junctions, decays, a multi-arc preference
    split, a binding arc capacity, a Storage and an (inflow-less) Reservoir."""
    set_default_pollutants()
    dates = [pd.Timestamp("2001-01-01") + pd.Timedelta(days=i) for i in range(T)]
    rng = np.random.default_rng(20261021)
    forcing = synthetic_forcing(dates, rng)

    def catch_data(seed, q):
        r = np.random.default_rng(seed)
        d = {}
        flow = q * (1 + 0.5 * np.sin(np.arange(T) / 30.0)) * r.lognormal(0, 0.3, T)
        for t, date in enumerate(dates):
            d[("flow", date)] = float(flow[t])
            for p in constants.POLLUTANTS:
                if p == "temperature":
                    d[(p, date)] = forcing[("temperature", date)] + 1.0
                elif p == "ph":
                    d[(p, date)] = 7.2
                elif p == "do":
                    d[(p, date)] = 0.009
                else:
                    d[(p, date)] = float(r.uniform(1e-5, 2e-3))
        return d

    tdata = {k: v for k, v in forcing.items() if k[0] == "temperature"}
    init = lambda v: {"volume": v, "do": 0.008, "org-phosphorus": v * 2e-5, "phosphate": v * 5e-5,  # noqa: E731
                      "ammonia": v * 1e-4, "solids": v * 1e-3, "bod": v * 1e-3, "cod": v * 2e-3, "ph": 7.1,
                      "temperature": 8.0, "nitrate": v * 2e-3, "nitrite": v * 1e-5, "org-nitrogen": v * 3e-4}
    decays = {"ammonia": {"constant": 0.03, "exponent": 1.01}, "bod": {"constant": 0.02, "exponent": 1.02}}
    nodes = [
        {"type_": "Catchment", "name": "c1", "data_input_dict": catch_data(1, 4e4)},
        {"type_": "Catchment", "name": "c2", "data_input_dict": catch_data(2, 1e4)},
        {"type_": "Catchment", "name": "c3", "data_input_dict": catch_data(3, 2e4)},
        {"type_": "River", "name": "r1", "length": 8000, "width": 12, "data_input_dict": tdata,
         "initial_storage": init(5e4)},
        {"type_": "River", "name": "r2", "length": 3000, "width": 8, "data_input_dict": tdata,
         "initial_storage": init(1e4), "decays": dict(decays)},
        {"type_": "River", "name": "r3", "length": 12000, "width": 20, "damp": 0.2, "data_input_dict": tdata,
         "initial_storage": init(9e4)},
        {"type_": "River", "name": "r4", "length": 5000, "width": 25, "data_input_dict": tdata,
         "initial_storage": 0},
        {"type_": "Node", "name": "j1"},
        {"type_": "Node", "name": "j2"},
        {"type_": "Groundwater", "name": "gw", "area": 1e6, "capacity": 3e5, "residence_time": 40.0,
         "initial_storage": init(1e5), "infiltration_threshold": 0.2, "infiltration_pct": 1e-6},
        {"type_": "Sewer", "name": "sw", "capacity": 500.0, "pipe_time": 1},
        {"type_": "Sewer", "name": "sw2", "capacity": 800.0},
        {"type_": "Storage", "name": "pond", "capacity": 2000.0, "area": 500.0, "initial_storage": init(300.0),
         "decays": {"nitrate": {"constant": 0.002, "exponent": 1.03}}, "data_input_dict": tdata},
        {"type_": "Reservoir", "name": "res", "capacity": 1e5, "area": 1e4, "initial_storage": init(4e4),
         "decays": dict(decays), "data_input_dict": tdata},
        {"type_": "Land", "name": "land", "data_input_dict": forcing, "surfaces": [
            {"type_": "ImperviousSurface", "surface": "urban", "area": 3e4, "pore_depth": 0.002,
             "pollutant_load": {"solids": 1e-5, "phosphate": 1e-8}},
            {"type_": "PerviousSurface", "surface": "rural", "area": 2e5, "depth": 0.6,
             "initial_storage": init(2e4), "decays": {"ammonia": {"constant": 0.05, "exponent": 1.005}}},
        ]},
        {"type_": "Waste", "name": "out"},
        {"type_": "Waste", "name": "out2"},
    ]
    arcs = [
        ("c1-r1", "c1", "r1", {}), ("c2-r2", "c2", "r2", {}), ("c3-j1", "c3", "j1", {}),
        ("r1-j1", "r1", "j1", {}), ("r2-r3", "r2", "r3", {}), ("j1-r3", "j1", "r3", {}),
        ("r3-j2", "r3", "j2", {}), ("j2-r4", "j2", "r4", {"preference": 2.0}),
        ("j2-out2", "j2", "out2", {"preference": 0.5, "capacity": 2.5e4}),
        ("r4-out", "r4", "out", {}),
        ("land-sw", "land", "sw", {}), ("land-gw", "land", "gw", {}),
        ("land-r1", "land", "r1", {"preference": 1.0}), ("land-r2", "land", "r2", {"preference": 3.0, "capacity": 900.0}),
        ("sw-sw2", "sw", "sw2", {"capacity": 300.0}), ("sw2-r3", "sw2", "r3", {}),
        ("gw-r3", "gw", "r3", {}), ("gw-sw", "gw", "sw", {}),
        ("pond-r4", "pond", "r4", {}), ("land-pond", "land", "pond", {}),
    ]
    m = Model()
    m.orchestration.insert(8, {"Storage": "distribute"})  # after Groundwater.distribute
    m.add_nodes(nodes)
    m.add_arcs([dict({"type_": "Arc", "name": a, "in_port": i, "out_port": o}, **kw) for a, i, o, kw in arcs])
    fm = flatten_ref(m, dates)
    ref = run_reference(m, dates)
    save("rivernet", fm, ref, "synthetic river/sewer network with decays, splits and binding capacities")


def scenario_reservoirs(T=400):
    """A RiverReservoir on a river (spill, environmental flow, a binding arc downstream so that some
    spill comes back) and the arc variants: a PullArc to an abstracting Reservoir, a PushArc to a
    Reservoir that also tries to abstract, a WeirArc and a SewerArc."""
    set_default_pollutants()
    dates = [pd.Timestamp("2003-01-01") + pd.Timedelta(days=i) for i in range(T)]
    rng = np.random.default_rng(20261022)
    forcing = synthetic_forcing(dates, rng)

    def catch_data(seed, q):
        r = np.random.default_rng(seed)
        d = {}
        flow = q * (1 + 0.8 * np.sin(np.arange(T) / 25.0)) * r.lognormal(0, 0.5, T)
        for t, date in enumerate(dates):
            d[("flow", date)] = float(flow[t])
            for p in constants.POLLUTANTS:
                if p == "temperature":
                    d[(p, date)] = forcing[("temperature", date)] + 1.0
                elif p == "ph":
                    d[(p, date)] = 7.2
                elif p == "do":
                    d[(p, date)] = 0.009
                else:
                    d[(p, date)] = float(r.uniform(1e-5, 2e-3))
        return d

    tdata = {k: v for k, v in forcing.items() if k[0] == "temperature"}
    init = lambda v: {"volume": v, "do": 0.008, "org-phosphorus": v * 2e-5, "phosphate": v * 5e-5,  # noqa: E731
                      "ammonia": v * 1e-4, "solids": v * 1e-3, "bod": v * 1e-3, "cod": v * 2e-3, "ph": 7.1,
                      "temperature": 8.0, "nitrate": v * 2e-3, "nitrite": v * 1e-5, "org-nitrogen": v * 3e-4}
    decays = {"ammonia": {"constant": 0.03, "exponent": 1.01}}
    nodes = [
        {"type_": "Catchment", "name": "c1", "data_input_dict": catch_data(1, 3e4)},
        {"type_": "Catchment", "name": "c2", "data_input_dict": catch_data(2, 8e3)},
        {"type_": "River", "name": "r1", "length": 6000, "width": 10, "data_input_dict": tdata, "initial_storage": init(4e4)},
        {"type_": "Node", "name": "j"},
        {"type_": "RiverReservoir", "name": "rr", "capacity": 2.5e5, "area": 4e4, "initial_storage": init(2.3e5),
         "environmental_flow": 9e3, "decays": dict(decays), "data_input_dict": tdata},
        {"type_": "Reservoir", "name": "farm", "capacity": 6e4, "area": 1e4, "initial_storage": init(1e4)},
        {"type_": "Reservoir", "name": "pond", "capacity": 3e4, "area": 1e4, "initial_storage": init(2e4)},
        {"type_": "River", "name": "r2", "length": 4000, "width": 15, "data_input_dict": tdata, "initial_storage": init(2e4)},
        {"type_": "Node", "name": "j2"},
        {"type_": "Waste", "name": "out"},
        {"type_": "Waste", "name": "out2"},
    ]
    arcs = [
        ("Arc", "c1-r1", "c1", "r1", {}), ("Arc", "r1-j", "r1", "j", {}),
        ("Arc", "j-rr", "j", "rr", {"preference": 2.0}),
        ("PullArc", "j-farm", "j", "farm", {"capacity": 4e3}),
        ("PushArc", "j-pond", "j", "pond", {"preference": 0.5, "capacity": 1.5e3}),
        ("WeirArc", "rr-j2", "rr", "j2", {}), ("Arc", "c2-j2", "c2", "j2", {}),
        ("Arc", "j2-r2", "j2", "r2", {"capacity": 2.2e4}), ("Arc", "j2-out2", "j2", "out2", {"capacity": 3e3, "preference": 0.2}),
        ("SewerArc", "r2-out", "r2", "out", {}),
    ]
    m = Model()
    m.orchestration.append({"Reservoir": "satisfy_environmental"})  # after the abstractions, before river routing
    m.add_nodes(nodes)
    m.add_arcs([dict({"type_": ty, "name": a, "in_port": i, "out_port": o}, **kw) for ty, a, i, o, kw in arcs])
    # only the RiverReservoir has satisfy_environmental: run it for that node alone, as a user would
    m.orchestration[-1] = {"RiverReservoirOnly": "satisfy_environmental"}
    m.nodes_type["RiverReservoirOnly"] = {"rr": m.nodes["rr"]}
    fm = flatten_ref(m, dates)
    ref = run_reference(m, dates)
    save("reservoirs", fm, ref, "RiverReservoir (spill, environmental flow) and PullArc / PushArc / WeirArc / SewerArc")


def scenario_irrigation(T=500):
    """A Land with an IrrigationSurface (pulls its crop water demand from a River, a Groundwater and a
    Reservoir through the Land's in-arcs, after the surfaces have run) and a VariableAreaSurface."""
    set_default_pollutants()
    dates = [pd.Timestamp("2004-01-01") + pd.Timedelta(days=i) for i in range(T)]
    rng = np.random.default_rng(20261023)
    forcing = synthetic_forcing(dates, rng)
    months = sorted({(d.year, d.month) for d in dates})
    monthly = {}
    for var in L.DEP_COLS:
        for (y, mth) in months:
            monthly[(var, pd.Timestamp(year=y, month=mth, day=1).to_period("M"))] = 1e-7
    tdata = {k: v for k, v in forcing.items() if k[0] == "temperature"}
    init = lambda v: {"volume": v, "do": 0.008, "org-phosphorus": v * 2e-5, "phosphate": v * 5e-5,  # noqa: E731
                      "ammonia": v * 1e-4, "solids": v * 1e-3, "bod": v * 1e-3, "cod": v * 2e-3, "ph": 7.1,
                      "temperature": 8.0, "nitrate": v * 2e-3, "nitrite": v * 1e-5, "org-nitrogen": v * 3e-4}

    def crop(kind, name, area, **extra):
        vol = area * 1.25 * 0.4 * 0.8
        soil = {"org-phosphorus": vol * 5e-3, "phosphate": vol * 5e-3, "ammonia": 0.0, "nitrate": vol,
                "nitrite": 0.0, "org-nitrogen": vol, "volume": vol}
        g = dict(GROWING_TEMPLATE)
        g.update(type_=kind, surface=name, area=area, pollutant_load={}, initial_storage=init(vol),
                 initial_soil_storage=soil, data_input_dict=monthly, rooting_depth=1.25,
                 decays={"ammonia": {"constant": 0.1, "exponent": 1.005}}, **extra)
        return g

    def catch_data(seed, q):
        r = np.random.default_rng(seed)
        d = {}
        flow = q * (1 + 0.5 * np.sin(np.arange(T) / 30.0)) * r.lognormal(0, 0.3, T)
        for t, date in enumerate(dates):
            d[("flow", date)] = float(flow[t])
            for p in constants.POLLUTANTS:
                d[(p, date)] = (forcing[("temperature", date)] + 1.0 if p == "temperature" else 7.2 if p == "ph"
                                else 0.009 if p == "do" else float(r.uniform(1e-5, 2e-3)))
        return d

    nodes = [
        {"type_": "Catchment", "name": "c1", "data_input_dict": catch_data(5, 2e3)},
        {"type_": "River", "name": "r1", "length": 4000, "width": 8, "data_input_dict": tdata, "initial_storage": init(8e3)},
        {"type_": "Reservoir", "name": "res", "capacity": 5e4, "area": 1e4, "initial_storage": init(3e4)},
        {"type_": "Land", "name": "farm", "data_input_dict": forcing, "surfaces": [
            crop("IrrigationSurface", "irrigated", 3e5, irrigation_coefficient=0.6),
            crop("VariableAreaSurface", "variable", 1e5, current_surface_area=6.5e4),
            {"type_": "PerviousSurface", "surface": "rural", "area": 5e4, "depth": 0.5, "initial_storage": init(5e3)},
        ]},
        {"type_": "Groundwater", "name": "gw", "area": 4e5, "capacity": 2e5, "residence_time": 60.0, "initial_storage": init(2e4)},
        {"type_": "Node", "name": "j"},
        {"type_": "Waste", "name": "out"},
    ]
    arcs = [
        ("c1-r1", "c1", "r1", {}), ("r1-j", "r1", "j", {}), ("j-out", "j", "out", {}),
        ("r1-farm", "r1", "farm", {"capacity": 900.0}), ("gw-farm", "gw", "farm", {"preference": 0.5}),
        ("res-farm", "res", "farm", {"preference": 0.2, "capacity": 150.0}),
        ("farm-gw", "farm", "gw", {}), ("farm-j", "farm", "j", {}), ("gw-j", "gw", "j", {}),
    ]
    m = Model()
    m.add_nodes(nodes)
    m.add_arcs([dict({"type_": "Arc", "name": a, "in_port": i, "out_port": o}, **kw) for a, i, o, kw in arcs])
    fm = flatten_ref(m, dates)
    ref = run_reference(m, dates)
    save("irrigation", fm, ref, "IrrigationSurface pulling through its Land (River / Groundwater / Reservoir) + VariableAreaSurface")


def scenario_distribution(T=300):
    """A plain Distribution with leakage (pulls pass through it enlarged by the leak, the leak goes to a
    Groundwater whose capacity binds now and then) supplying a base Demand (constant demand with its own
    loads) and a ResidentialDemand from an FWTW fed by a Reservoir; the used water leaves through a Sewer."""
    set_default_pollutants()
    dates = [pd.Timestamp("2005-01-01") + pd.Timedelta(days=i) for i in range(T)]
    rng = np.random.default_rng(20261024)
    forcing = synthetic_forcing(dates, rng)
    tdata = {k: v for k, v in forcing.items() if k[0] == "temperature"}
    init = lambda v: {"volume": v, "do": 0.008, "org-phosphorus": v * 2e-5, "phosphate": v * 5e-5,  # noqa: E731
                      "ammonia": v * 1e-4, "solids": v * 1e-3, "bod": v * 1e-3, "cod": v * 2e-3, "ph": 7.1,
                      "temperature": 9.0, "nitrate": v * 2e-3, "nitrite": v * 1e-5, "org-nitrogen": v * 3e-4}
    house = {"do": 0.004, "org-phosphorus": 2e-7, "phosphate": 5e-7, "ammonia": 3e-6, "solids": 1e-5, "bod": 2e-5,
             "cod": 4e-5, "ph": 7.4, "temperature": 14.0, "nitrate": 1e-6, "nitrite": 1e-8, "org-nitrogen": 2e-6}
    nodes = [
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
        {"type_": "Waste", "name": "out"},
    ]
    arcs = [
        ("res-fwtw", "res", "fwtw", {}), ("fwtw-mains", "fwtw", "mains", {}), ("fwtw-foul", "fwtw", "foul", {}),
        ("mains-works", "mains", "works", {"capacity": 640.0}), ("mains-houses", "mains", "houses", {}),
        ("mains-gw", "mains", "gw", {}), ("works-foul", "works", "foul", {}), ("houses-foul", "houses", "foul", {}),
        ("foul-j", "foul", "j", {}), ("gw-j", "gw", "j", {}), ("j-out", "j", "out", {}),
    ]
    m = Model()
    m.add_nodes(nodes)
    m.add_arcs([dict({"type_": "Arc", "name": a, "in_port": i, "out_port": o}, **kw) for a, i, o, kw in arcs])
    fm = flatten_ref(m, dates)
    ref = run_reference(m, dates)
    save("distribution", fm, ref, "Distribution with leakage, base Demand with a constant demand, ResidentialDemand, FWTW")


def scenario_rivertree(T=380):
    """Dendritic river network (the network engine's case): catchments and rivers joined through
    junctions (also chained and shared ones), a binding river-arc capacity, decays, a dry
    headwater river and a Waste outlet."""
    set_default_pollutants()
    dates = [pd.Timestamp("2003-06-01") + pd.Timedelta(days=i) for i in range(T)]
    rng = np.random.default_rng(20261101)
    forcing = synthetic_forcing(dates, rng)
    tdata = {k: v for k, v in forcing.items() if k[0] == "temperature"}

    def catch_data(seed, q):
        r = np.random.default_rng(seed)
        d = {}
        flow = q * (1 + 0.6 * np.sin(np.arange(T) / 25.0 + seed)) * r.lognormal(0, 0.4, T)
        for t, date in enumerate(dates):
            d[("flow", date)] = float(flow[t])
            for p in constants.POLLUTANTS:
                if p == "temperature":
                    d[(p, date)] = forcing[("temperature", date)] + 0.5 * seed
                elif p == "ph":
                    d[(p, date)] = 7.0 + 0.05 * seed
                elif p == "do":
                    d[(p, date)] = 0.009
                else:
                    d[(p, date)] = float(r.uniform(1e-5, 2e-3))
        return d

    init = lambda v: {"volume": v, "do": 0.008, "org-phosphorus": v * 2e-5, "phosphate": v * 5e-5,  # noqa: E731
                      "ammonia": v * 1e-4, "solids": v * 1e-3, "bod": v * 1e-3, "cod": v * 2e-3, "ph": 7.1,
                      "temperature": 9.0, "nitrate": v * 2e-3, "nitrite": v * 1e-5, "org-nitrogen": v * 3e-4}
    river = lambda name, length, width, vol, **kw: dict(  # noqa: E731
        {"type_": "River", "name": name, "length": length, "width": width, "data_input_dict": tdata,
         "initial_storage": init(vol) if vol else 0}, **kw)
    nodes = [{"type_": "Catchment", "name": "c%d" % i, "data_input_dict": catch_data(i, q)}
             for i, q in enumerate([3e4, 1e4, 2e4, 5e3, 8e3, 1.5e4, 4e3], start=1)]
    nodes += [
        river("r1", 6000, 10, 4e4), river("r2", 2500, 6, 8e3), river("r3", 9000, 15, 7e4, damp=0.15),
        river("r4", 7000, 22, 9e4), river("r5", 3000, 7, 0),
        river("r6", 11000, 30, 1.5e5, decays={"ammonia": {"constant": 0.03, "exponent": 1.01}}),
        river("r7", 4000, 9, 2e4), river("r8", 3500, 8, 1e4,
                                         decays={"bod": {"constant": 0.02, "exponent": 1.02},
                                                 "nitrite": {"constant": 0.01, "exponent": 1.0}}),
        river("r9", 1500, 4, 0),
        {"type_": "Node", "name": "j1"}, {"type_": "Node", "name": "j2"}, {"type_": "Node", "name": "j3"},
        {"type_": "Node", "name": "j4"},
        {"type_": "Waste", "name": "outlet"},
    ]
    arcs = [
        ("c1-r1", "c1", "r1", {}), ("c2-r2", "c2", "r2", {}), ("r1-j1", "r1", "j1", {}), ("r2-j1", "r2", "j1", {}),
        ("j1-r3", "j1", "r3", {}), ("c3-r3", "c3", "r3", {}), ("r3-r4", "r3", "r4", {"capacity": 4.5e4}),
        ("c4-j2", "c4", "j2", {}), ("j2-j3", "j2", "j3", {}), ("j3-r5", "j3", "r5", {}), ("r5-r4", "r5", "r4", {}),
        ("r4-r6", "r4", "r6", {}), ("c5-r7", "c5", "r7", {"capacity": 9e3}), ("r7-r6", "r7", "r6", {}),
        ("c6-r8", "c6", "r8", {}), ("r8-r6", "r8", "r6", {}), ("r9-r7", "r9", "r7", {}),
        ("r6-j4", "r6", "j4", {}), ("c7-j4", "c7", "j4", {}), ("j4-outlet", "j4", "outlet", {}),
    ]
    m = Model()
    m.add_nodes(nodes)
    m.add_arcs([dict({"type_": "Arc", "name": a, "in_port": i, "out_port": o}, **kw) for a, i, o, kw in arcs])
    fm = flatten_ref(m, dates)
    ref = run_reference(m, dates)
    save("rivertree", fm, ref, "dendritic river network with junction chains, binding capacities and decays")


def _config3_setup():
    df = pd.read_csv(os.path.join(REF, "docs/demo/data/processed/timeseries_data.csv"))
    df.loc[df.variable == "flow", "value"] *= constants.M3_S_TO_M3_DT
    df.loc[df.variable == "precipitation", "value"] *= constants.MM_TO_M
    df.date = pd.to_datetime(df.date)
    did = {site: {(v, d): x for v, d, x in zip(g.variable, g.date, g.value)} for site, g in df.groupby("site")}
    dates = sorted(pd.Timestamp(x) for x in df.date.unique())
    pol = [v for v in df.variable.unique().tolist() if v not in ("flow", "precipitation", "et0")]
    constants.POLLUTANTS = pol
    constants.NON_ADDITIVE_POLLUTANTS = ["temperature"]
    constants.ADDITIVE_POLLUTANTS = [p for p in pol if p != "temperature"]
    constants.FLOAT_ACCURACY = 1e-8
    return did, dates


def _config3_member(did, f):
    """One member of SURVEY 8d config 3 with the six parameter factors ``f`` (pervious area, field
    capacity, depth, impervious area, groundwater area, groundwater residence time)."""
    from wsimod.nodes.catchment import Catchment
    from wsimod.nodes.land import Land
    from wsimod.nodes.nodes import Node
    from wsimod.nodes.storage import Groundwater
    from wsimod.nodes.waste import Waste
    from wsimod.arcs.arcs import Arc

    deposition = {"boron": 100e-10, "calcium": 70e-7, "chloride": 60e-10, "fluoride": 0.2e-7, "magnesium": 6e-7,
                  "nitrate": 2e-9, "nitrogen": 4e-7, "potassium": 7e-7, "silicon": 7e-9, "sodium": 30e-9,
                  "sulphate": 70e-7}
    nodes = [Waste(name="thames_above_abingdon")]
    nodes += [Catchment(name=c, data_input_dict=did[c]) for c in ("evenlode", "thames", "ray", "cherwell")]
    nodes.append(Land(name="oxford_land", data_input_dict=did["oxford_land"], surfaces=[
        {"type_": "PerviousSurface", "area": 2e7 * f[0], "pollutant_load": dict(deposition), "surface": "rural",
         "field_capacity": 0.3 * f[1], "depth": 0.5 * f[2], "initial_storage": 2e7 * f[0] * 0.4 * 0.5 * f[2]},
        {"type_": "ImperviousSurface", "area": 1e7 * f[3], "pollutant_load": dict(deposition),
         "surface": "urban", "initial_storage": 5e6 * f[3]}]))
    nodes.append(Groundwater(capacity=3.2e9, area=3.2e8 * f[4], name="gw", residence_time=20 * f[5]))
    nodes += [Node(name=n) for n in ("farmoor_abstraction", "evenlode_thames", "cherwell_ray",
                                     "cherwell_thames", "thames_mixer")]
    by = {n.name: n for n in nodes}
    arcs = [Arc(in_port=by[i], out_port=by[o], name=a) for a, i, o in [
        ("evenlode_to_thames", "evenlode", "evenlode_thames"), ("thames_to_thames", "thames", "evenlode_thames"),
        ("ray_to_cherwell", "ray", "cherwell_ray"), ("cherwell_to_cherwell", "cherwell", "cherwell_ray"),
        ("thames_to_farmoor", "evenlode_thames", "farmoor_abstraction"),
        ("farmoor_to_mixer", "farmoor_abstraction", "thames_mixer"),
        ("cherwell_to_mixer", "cherwell_ray", "thames_mixer"), ("land_to_gw", "oxford_land", "gw"),
        ("land_to_mixer", "oxford_land", "thames_mixer"), ("gw_to_mixer", "gw", "thames_mixer"),
        ("mixer_to_waste", "thames_mixer", "thames_above_abingdon")]]
    m = Model()
    m.add_instantiated_nodes(nodes)
    m.add_instantiated_arcs(arcs)
    return m


def scenario_config3(R=2):
    """the Oxford catchment's land + river-side nodes (Land{Pervious 2e7 m2,
    Impervious 1e7 m2}, Groundwater, the junction Nodes, Waste, the 4 Catchment inflows as
    prescribed sources) on the demo's own time series (1456 days, the 12 Oxford pollutants,
    temperature the only non-additive one), R members with parameters x log-uniform(0.8, 1.25)."""
    from wsi_b200.ensemble import config3_factors

    did, dates = _config3_setup()
    factors = config3_factors(R)
    fms, refs = [], []
    for member in range(R):
        m = _config3_member(did, factors[member])
        fms.append(flatten_ref(m, dates))
        refs.append(run_reference(m, dates))
    fm = merge_replicas(fms)
    ref = {k: np.stack([r[k] for r in refs], axis=-1) for k in refs[0] if k != "n_advisories"}
    ref["n_advisories"] = np.array([int(sum(r["n_advisories"][0] for r in refs))])
    save("config3", fm, ref, "SURVEY 8d config 3: Oxford land + river-side subset, %d members, %d days" % (R, len(dates)))
    set_default_pollutants()


def scenario_config3_members(R=16):
    """The first 16 members of the config-3 ensemble (wsi_b200.ensemble.config3_factors), every one
    built and run with the reference; to keep the fixture small only three arcs are stored: what the
    Land sends to the groundwater and to the river, and the catchment outlet (which sums everything)."""
    from wsi_b200.ensemble import config3_factors

    did, dates = _config3_setup()
    factors = config3_factors(R)
    keep = ["land_to_gw", "land_to_mixer", "mixer_to_waste"]
    fms, flows, adv = [], [], 0
    for member in range(R):
        m = _config3_member(did, factors[member])
        fms.append(flatten_ref(m, dates))
        r = run_reference(m, dates)
        idx = [list(m.arcs.keys()).index(a) for a in keep]
        flows.append(r["flows"][:, idx])
        adv += int(r["n_advisories"][0])
    fm = merge_replicas(fms)
    ref = {"flows_kept": np.stack(flows, axis=-1), "kept_arcs": np.array(keep), "n_advisories": np.array([adv])}
    save("config3_members16", fm, ref, "config 3: %d reference-built members, 3 arcs, %d days" % (R, len(dates)))
    set_default_pollutants()


def scenario_oxford():
    """BASELINE.json configs[0], second half: the reference's own Oxford demo
    (wsimod/demo/create_oxford.py:149-394) -- 18 nodes, 21 arcs, 12 Oxford pollutants, 1456 days:
    four Catchments, junctions, Land, Groundwater, combined Sewer with a pipe time-area diagram and
    a preference split between WWTW and overflow, WWTW, FWTW fed from the Farmoor Reservoir, the
    reservoir's abstraction through a capacity-limited arc, ResidentialDemand drawing through a
    distribution junction."""
    from wsimod.demo.create_oxford import create_oxford_model

    m = create_oxford_model(os.path.join(REF, "docs", "demo", "data"))
    dates = m.dates
    fm = flatten_ref(m, dates)
    ref = run_reference(m, dates)
    save("oxford", fm, ref, "create_oxford_model, %d days" % len(dates))
    set_default_pollutants()


SCENARIOS = {
    "quickstart": scenario_quickstart,
    "quickstart_simple": scenario_quickstart_simple,
    "quickstart_p5": scenario_quickstart_p5,
    "quickstart_p14": scenario_quickstart_p14,
    "fixture3607": scenario_fixture3607,
    "config2": scenario_config2,
    "rivernet": scenario_rivernet,
    "reservoirs": scenario_reservoirs,
    "irrigation": scenario_irrigation,
    "distribution": scenario_distribution,
    "rivertree": scenario_rivertree,
    "config3": scenario_config3,
    "hourly": scenario_hourly,
    "oxford": scenario_oxford,
    "queuegw": scenario_queuegw,
    "city": scenario_city,
    "fixture3607_full": scenario_fixture3607_full,
    "config3_members16": scenario_config3_members,
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(SCENARIOS)
    for n in names:
        SCENARIOS[n]()
