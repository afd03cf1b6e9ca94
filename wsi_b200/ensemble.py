"""Ensembles of independent catchments: one template model replicated ``R`` times with
per-member parameters, flattened once and stepped by one kernel launch per timestep.

The reference has no ensemble facility (each member would be its own ``Model`` object and
Python loop); here members are the fastest-varying index of every SoA plane, which is what
makes the kernels' loads coalesce.  ``config2_*`` build the benchmark workload named by
BASELINE.json ``configs[1]`` and SURVEY 8d "Config 2".
"""
from __future__ import annotations

import numpy as np

from .backend import layout as L
from .backend.flatten import FlatModel
from .core import constants

# "Spring wheat" GrowingSurface of the reference's real-catchment fixture
# (tests/test_model_data.zip:config.yml)
GROWING_TEMPLATE = dict(
    surface_coefficient=0.05, field_capacity=0.3577337861061096, wilting_point=0.14462512731552124,
    crop_factor_stages=[0.0, 0.0, 0.3, 0.3, 1.15, 1.15, 0.325, 0.0, 0.0],
    crop_factor_stage_dates=[0, 73, 74, 121, 166, 196, 274, 275, 366],
    ihacres_p=10.0, rooting_depth=1.25, percolation_coefficient=0.25, total_porosity=0.4,
    ET_depletion_factor=0.55, sowing_day=74, harvest_day=274, infiltration_capacity=0.06,
)


def daily_dates(T, start="2000-01-01"):
    import pandas as pd

    t0 = pd.Timestamp(start)
    return [t0 + pd.Timedelta(days=i) for i in range(T)]


def synthetic_forcing_arrays(dates, n_series, seed=20261019):
    """[T, n_series, 3] precipitation (m), et0 (m), temperature (C): SURVEY 8d config 2."""
    rng = np.random.default_rng(seed)
    doy = np.array([d.dayofyear for d in dates], dtype=np.float64)[:, None]
    T = len(dates)
    temperature = 10 + 8 * np.sin(2 * np.pi * (doy - 110) / 365) + rng.normal(0, 2, (T, n_series))
    et0 = np.maximum(2e-4, 1e-3 * (1.5 + 1.2 * np.sin(2 * np.pi * (doy - 100) / 365))) * np.ones((1, n_series))
    wet = rng.random((T, n_series)) < 0.45
    precipitation = wet * rng.gamma(0.8, 6e-3, (T, n_series))
    return np.ascontiguousarray(np.stack([precipitation, et0, temperature], axis=2))


def config2_spec(dates, forcing_col, scale=1.0, area_scale=1.0):
    """Node and arc specifications (the ``add_nodes`` / ``add_arcs`` dictionaries of
    ``orchestration/model.py:642-778``) of one config-2 catchment, {Land(Impervious, Pervious,
    Growing) -> Sewer, Groundwater, junction Node -> Waste}.  ``forcing_col`` is ``[T, 3]``.
    The same dictionaries build the member with this package's ``Model`` and with the reference's
    (bench.py's reference arm, baseline/reference_arm.py); ``scale`` / ``area_scale`` perturb an
    object-level member the way tests/golden/make_golden.py does."""
    forcing = {}
    for t, d in enumerate(dates):
        forcing[("precipitation", d)] = float(forcing_col[t, 0])
        forcing[("et0", d)] = float(forcing_col[t, 1])
        forcing[("temperature", d)] = float(forcing_col[t, 2])
    months = []
    for d in dates:
        m = d.to_period("M")
        if not months or months[-1] != m:
            months.append(m)
    monthly = {(var, m): 1e-7 for var in L.DEP_COLS for m in months}  # kg/m2 as in land_demo.py
    a_imp, a_perv, a_grow = 2.0e4 * area_scale, 5.0e4 * area_scale, 2.0e5 * area_scale
    vol = a_grow * 1.25 * 0.4 * 0.9
    init = {"do": 0.01, "org-phosphorus": vol * 1e-4, "phosphate": vol * 1e-4, "ammonia": vol * 5e-5,
            "solids": vol * 1e-2, "bod": vol * 1e-3, "cod": vol * 2e-2, "ph": 7.0, "temperature": 11.0,
            "nitrate": vol * 4e-3, "nitrite": vol * 2e-5, "org-nitrogen": vol * 5e-3, "volume": vol}
    soil = {"org-phosphorus": vol * 5e-3, "phosphate": vol * 5e-3, "ammonia": 0.0, "nitrate": vol, "nitrite": 0.0,
            "org-nitrogen": vol, "volume": vol}
    grow = dict(GROWING_TEMPLATE)
    grow.update(type_="GrowingSurface", surface="crop", area=a_grow, pollutant_load={}, initial_storage=init,
                initial_soil_storage=soil, data_input_dict=monthly,
                decays={"ammonia": {"constant": 0.1 * scale, "exponent": 1.005},
                        "nitrite": {"constant": 0.05 * scale, "exponent": 1.005}},
                infiltration_capacity=0.06 * scale)
    imp = {"type_": "ImperviousSurface", "surface": "urban", "area": a_imp, "pore_depth": 0.004 * scale,
           "pollutant_load": {"phosphate": 1e-10, "solids": 1e-5, "bod": 1e-5, "ammonia": 1e-9, "temperature": 13.0,
                              "do": 0.005, "ph": 7.5}}
    perv = {"type_": "PerviousSurface", "surface": "rural", "area": a_perv, "depth": 0.5,
            "pollutant_load": {"phosphate": 1e-7}, "initial_storage": a_perv * 0.5 * 0.4 * 0.5,
            "infiltration_capacity": 0.05}
    land = {"type_": "Land", "name": "land", "data_input_dict": forcing, "surfaces": [imp, perv, grow],
            "surface_residence_time": 2.0 * scale, "subsurface_residence_time": 10.0,
            "percolation_residence_time": 30.0 / scale}
    sewer = {"type_": "Sewer", "name": "sewer", "capacity": a_imp * 0.01 * scale}
    gw = {"type_": "Groundwater", "name": "gw", "area": a_perv + a_grow, "capacity": (a_perv + a_grow) * 0.5,
          "residence_time": 100.0 * scale, "initial_storage": (a_perv + a_grow) * 0.01}
    arcs = [("land-sewer", "land", "sewer"), ("land-gw", "land", "gw"), ("land-river", "land", "river"),
            ("sewer-river", "sewer", "river"), ("gw-river", "gw", "river"), ("river-outlet", "river", "outlet")]
    nodes = [sewer, land, gw, {"type_": "Node", "name": "river"}, {"type_": "Waste", "name": "outlet"}]
    return nodes, [{"type_": "Arc", "name": a, "in_port": i, "out_port": o} for a, i, o in arcs]


def config2_template_model(dates, forcing_col):
    """The config-2 catchment built through the drop-in ``Model`` API."""
    from .orchestration.model import Model

    constants.set_default_pollutants()
    constants.FLOAT_ACCURACY = 1e-11
    nodes, arcs = config2_spec(dates, forcing_col)
    m = Model()
    m.add_nodes(nodes)
    m.add_arcs(arcs)
    m.dates = dates
    return m


def _scale(fm, group, name, factor, objs=None, n=None):
    """Multiply field planes of the listed objects by a per-replica factor ``[R]``."""
    v = fm.view(group, name, n)  # [n_planes, n_obj, R]
    objs = range(v.shape[1]) if objs is None else objs
    for o in objs:
        v[:, o, :] *= factor[None, :]


def perturb_config2(fm, seed=20261019):
    """Per-member parameters: fixture values x log-uniform(0.8, 1.25), areas log-uniform
    (1e4, 1e7) m2 for the largest surface.  Extensive quantities
    (capacities, initial storages, nutrient pools) scale with area; derived parameters are
    kept consistent with the ones they derive from."""
    rng = np.random.default_rng(seed)
    R = fm.R
    lu = lambda: np.exp(rng.uniform(np.log(0.8), np.log(1.25), R))  # noqa: E731
    area0 = fm.view("surf", "area")[0, :, 0].max()
    a = np.exp(rng.uniform(np.log(1e4), np.log(1e7), R)) / area0
    V = fm.V
    add_planes = [0] + [fm.pollutants.index(p) + 1 for p in fm.additive]
    # extensive scaling with area
    for f in ("area", "capacity"):
        _scale(fm, "surf", f, a)
    st = fm.view("surf", "storage")
    st[add_planes] *= a[None, None, :]
    for f in ("pool_fast", "pool_humus", "pool_dis_inorg", "pool_dis_org", "pool_ads_inorg"):
        _scale(fm, "surf", f, a)
    for f in ("capacity", "area"):
        _scale(fm, "tank", f, a)
    for f in ("storage", "storage_"):
        fm.view("tank", f)[add_planes] *= a[None, None, :]
    _scale(fm, "sewer", "capacity", a * lu())
    # intensive parameters
    for f in ("res_surface", "res_subsurface", "res_percolation"):
        _scale(fm, "land", f, lu())
    _scale(fm, "tank", "residence_time", lu())
    for f in ("decay_const", "infiltration_capacity", "et0_to_e", "surface_coefficient", "uptake1", "uptake3",
              "erodibility", "denpar", "kfr", "kadsdes", "degrhpar", "dishpar", "minfpar", "disfpar", "immobdpar",
              "hsatINs", "cohesion"):
        _scale(fm, "surf", f, lu())
    pc = fm.view("surf", "percolation_coefficient")
    pc *= lu()[None, None, :]
    np.clip(pc, 0.0, 0.95, out=pc)
    fm.view("surf", "subsurface_coefficient")[:] = 1 - pc
    # impervious pore depth: capacity of surfaces of type Impervious
    types = fm.view("surf", "type")[0, :, 0]
    imp = [i for i, t in enumerate(types) if int(t) == L.ST["ImperviousSurface"]]
    _scale(fm, "surf", "capacity", lu(), objs=imp)
    return fm


def build_config2(R, T, seed=20261019, members_per_series=100, template=None):
    """Flat model of the config-2 workload: ``R`` members, ``T`` daily steps from 2000-01-01,
    one forcing series per ``members_per_series`` members."""
    dates = daily_dates(T)
    n_series = max(1, (R + members_per_series - 1) // members_per_series)
    forcing = synthetic_forcing_arrays(dates, n_series, seed)
    if template is None:
        template = config2_template_model(dates, forcing[:, 0, :]).flatten()
    fm = template.replicate(R)
    fm.node_forcing = forcing
    fid = (np.arange(R) // members_per_series).astype(np.float64)
    fm.view("land", "forcing_id")[0, :, :] = fid[None, :]
    perturb_config2(fm, seed)
    return fm


def land_out_arcs(fm):
    """Arc indices leaving Land nodes (the recorded outputs O of SURVEY 8d)."""
    land = set(np.nonzero(fm.node_type == L.NT["Land"])[0].tolist())
    return [a for a in range(len(fm.arc_names)) if int(fm.arc_in[a]) in land]


def algorithmic_bytes_per_replica_step(fm, n_recorded_planes):
    """SURVEY 8d field-count formula ``8 * (2 S + Q + F + O)`` evaluated on the actual model:
    S persistent state doubles (read and written once per step), Q parameters read once, F
    forcing doubles, O recorded outputs.  Counts only planes the step semantically needs --
    never allocated, padded or scratch planes."""
    V = fm.V
    types = fm.view("surf", "type")[0, :, 0].astype(int)
    n_decay = fm.view("surf", "n_decay")[0, :, 0].astype(int)
    has_load = fm.view("surf", "has_load")[0, :, 0].astype(int)
    has_dep = fm.view("surf", "has_dep")[0, :, 0].astype(int)
    nitrate = "nitrate" in fm.pollutants
    S = Q = F = 0
    for ty, nd, hl, hd in zip(types, n_decay, has_load, has_dep):
        S += V + 2 * nd  # storage + total_decayed + pre_decay
        Q += 3 + 3 * nd  # type, area, capacity + decay (pol, const, exp)
        Q += fm.P if hl else 0
        F += 6 if hd else 0
        if ty == L.ST["ImperviousSurface"]:
            Q += 1
        if ty >= L.ST["PerviousSurface"]:
            Q += 13  # depth, fc_m, wp_m, inf cap, 3 coefficients, ihacres_p, et0 coef, 4 soil temperature
        if ty == L.ST["GrowingSurface"]:
            ns = int(fm.view("surf", "n_stage")[0, 0, 0])
            S += 4 + (10 if nitrate else 0)  # crop state + pools
            Q += 7 + 2 * ns + 15  # calendar scalars, stage tables, 3 x 5 calendar tables
            if nitrate:
                Q += 24 + 16  # process constants + nutrient-pool rates/fractions
                F += 6  # fertiliser + manure
    n_land = fm.n_obj["land"]
    S += n_land * 3 * V
    Q += n_land * 4
    F += n_land * 3
    n_tank, n_sewer = fm.n_obj["tank"], fm.n_obj["sewer"]
    S += n_tank * V
    Q += n_tank * 5
    S += n_sewer * 2 * V  # storage + active_storage (queue slots are empty for pipe_time 0)
    Q += n_sewer * 2
    Q += 2 * len(fm.arc_names)
    O = n_recorded_planes
    return 8 * (2 * S + Q + F + O), dict(S=S, Q=Q, F=F, O=O)


def config3_factors(R, seed=20261020):
    """``[R, 6]`` parameter factors of the config-3 ensemble: member 0 is the Oxford
    demo's own parameter set, every other member draws log-uniform(0.8, 1.25) factors for the
    pervious area, field capacity, soil depth, impervious area, groundwater area and groundwater
    residence time from ``default_rng(20261020)``, six per member in member order."""
    rng = np.random.default_rng(seed)
    f = np.ones((int(R), 6))
    if R > 1:
        f[1:] = np.exp(rng.uniform(np.log(0.8), np.log(1.25), (int(R) - 1, 6)))
    return f


def build_config3(R, template, factors=None):
    """``R`` distinct members of the Oxford catchment's land + river-side subset.

    ``template`` is a flat model whose member 0 is the Oxford demo's own parameter set (the golden
    fixture ``tests/golden/config3.npz``, generated from the reference's Oxford data, which is not
    shipped here).  Member ``i`` is that member with the factors ``config3_factors(R)[i]`` applied to
    the flat planes exactly as the constructors of ``nodes/land.py`` / ``nodes/storage.py`` derive
    them (``land.py:794-814``: field_capacity_m = field_capacity * depth, the tank's depth =
    depth * total_porosity, capacity = area * depth; ``tests/golden/config3_members16.npz`` holds
    the same members built object by object with the reference).  Forcing and the four catchment
    inflow series are shared by all members (broadcast loads)."""
    R = int(R)
    f = config3_factors(R) if factors is None else np.asarray(factors, dtype=np.float64)
    fm = template.take_replicas(np.zeros(R, dtype=np.int64))
    types = fm.view("surf", "type")[0, :, 0].astype(int)
    perv = [i for i, t in enumerate(types) if t == L.ST["PerviousSurface"]]
    imp = [i for i, t in enumerate(types) if t == L.ST["ImperviousSurface"]]

    def mul(group, name, objs, factor, planes=None):
        v = fm.view(group, name)
        rows = range(v.shape[0]) if planes is None else planes
        for o in objs:
            for p in rows:
                v[p, o, :] *= factor

    mul("surf", "area", perv, f[:, 0])
    mul("surf", "capacity", perv, f[:, 0] * f[:, 2])
    mul("surf", "depth", perv, f[:, 2])
    mul("surf", "field_capacity_m", perv, f[:, 1] * f[:, 2])
    mul("surf", "wilting_point_m", perv, f[:, 2])
    mul("surf", "storage", perv, f[:, 0] * f[:, 2], planes=[0])
    mul("surf", "area", imp, f[:, 3])
    mul("surf", "capacity", imp, f[:, 3])
    mul("surf", "storage", imp, f[:, 3], planes=[0])
    gw = [int(fm.node_gidx[i]) for i, t in enumerate(fm.node_type) if int(t) == L.NT["Groundwater"]]
    mul("tank", "area", gw, f[:, 4])
    mul("tank", "residence_time", gw, f[:, 5])
    return fm
