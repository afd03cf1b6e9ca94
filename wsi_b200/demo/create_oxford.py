"""The Oxford demonstration catchment as a ``wsi_b200`` model (reference:
wsimod/demo/create_oxford.py:149-394 -- same nodes, arcs, parameters and input file layout, so
scripts that call ``create_oxford_model(data_folder)`` keep working).  18 nodes and 21 arcs: four
river catchments, the junctions between them, the Farmoor reservoir and its abstraction, fresh- and
waste-water treatment works, Oxford's residential demand, a combined sewer, land and groundwater."""
from __future__ import annotations

import os

from ..arcs.arcs import Arc
from ..core import constants
from ..nodes import (FWTW, WWTW, Catchment, Groundwater, Land, Node, Reservoir, ResidentialDemand, Sewer, Waste)
from ..orchestration.model import Model

# kg/m2/timestep deposited on both land surfaces
DEPOSITION = {"boron": 100e-10, "calcium": 70e-7, "chloride": 60e-10, "fluoride": 0.2e-7, "magnesium": 6e-7,
              "nitrate": 2e-9, "nitrogen": 4e-7, "potassium": 7e-7, "silicon": 7e-9, "sodium": 30e-9,
              "sulphate": 70e-7}
# household wastewater: (concentration, unit) per determinand, times the per-capita consumption
HOUSEHOLD = {"boron": (500, "ug"), "calcium": (150, "mg"), "chloride": (180, "mg"), "fluoride": (0.4, "mg"),
             "magnesium": (30, "mg"), "nitrate": (60, "mg"), "nitrogen": (50, "mg"), "potassium": (30, "mg"),
             "silicon": (20, "mg"), "sodium": (200, "mg"), "sulphate": (250, "mg")}
ARCS = [  # (name, from, to, extra)
    ("evenlode_to_thames", "evenlode", "evenlode_thames", {}),
    ("thames_to_thames", "thames", "evenlode_thames", {}),
    ("ray_to_cherwell", "ray", "cherwell_ray", {}),
    ("cherwell_to_cherwell", "cherwell", "cherwell_ray", {}),
    ("thames_to_farmoor", "evenlode_thames", "farmoor_abstraction", {}),
    ("farmoor_to_mixer", "farmoor_abstraction", "thames_mixer", {}),
    ("cherwell_to_mixer", "cherwell_ray", "thames_mixer", {}),
    ("wwtw_to_mixer", "oxford_wwtw", "thames_mixer", {}),
    ("sewer_overflow", "combined_sewer", "thames_mixer", {"preference": 1e-10}),
    ("mixer_to_waste", "thames_mixer", "thames_above_abingdon", {}),
    ("abstraction_to_farmoor", "farmoor_abstraction", "farmoor", {"capacity": 5e4}),
    ("distribution_to_demand", "oxford_distribution", "oxford", {}),
    ("demand_to_sewer", "oxford", "combined_sewer", {}),
    ("land_to_sewer", "oxford_land", "combined_sewer", {}),
    ("sewer_to_wwtw", "combined_sewer", "oxford_wwtw", {"preference": 1e10}),
    ("fwtw_to_sewer", "oxford_fwtw", "combined_sewer", {}),
    ("fwtw_to_distribution", "oxford_fwtw", "oxford_distribution", {}),
    ("reservoir_to_fwtw", "farmoor", "oxford_fwtw", {}),
    ("land_to_gw", "oxford_land", "gw", {}),
    ("garden_to_gw", "oxford", "gw", {}),
    ("gw_to_mixer", "gw", "thames_mixer", {}),
]
CREATION_ORDER = [
    "fwtw_to_distribution", "abstraction_to_farmoor", "sewer_to_wwtw", "sewer_overflow", "evenlode_to_thames",
    "thames_to_thames", "ray_to_cherwell", "cherwell_to_cherwell", "thames_to_farmoor", "farmoor_to_mixer",
    "cherwell_to_mixer", "wwtw_to_mixer", "mixer_to_waste", "distribution_to_demand", "reservoir_to_fwtw",
    "fwtw_to_sewer", "demand_to_sewer", "land_to_sewer", "land_to_gw", "garden_to_gw", "gw_to_mixer",
]


def create_oxford_model(data_folder):
    import pandas as pd

    df = pd.read_csv(os.path.join(data_folder, "processed", "timeseries_data.csv"))
    df.loc[df.variable == "flow", "value"] *= constants.M3_S_TO_M3_DT
    df.loc[df.variable == "precipitation", "value"] *= constants.MM_TO_M
    df.date = pd.to_datetime(df.date)
    inputs = {site: {(v, d): x for v, d, x in zip(g.variable, g.date, g.value)} for site, g in df.groupby("site")}
    dates = [pd.Timestamp(x) for x in sorted(df.date.unique())]
    constants.POLLUTANTS = [v for v in df.variable.unique().tolist() if v not in ("flow", "precipitation", "et0")]
    constants.NON_ADDITIVE_POLLUTANTS = ["temperature"]
    constants.ADDITIVE_POLLUTANTS = list(set(constants.POLLUTANTS).difference(constants.NON_ADDITIVE_POLLUTANTS))
    constants.FLOAT_ACCURACY = 1e-8

    per_capita = 0.15
    unit = {"ug": constants.UG_L_TO_KG_M3, "mg": constants.MG_L_TO_KG_M3}
    household = {k: c * unit[u] * per_capita for k, (c, u) in HOUSEHOLD.items()}
    household["temperature"] = 14
    surfaces = [
        {"type_": "PerviousSurface", "area": 2e7, "pollutant_load": DEPOSITION, "surface": "rural",
         "field_capacity": 0.3, "depth": 0.5, "initial_storage": 2e7 * 0.4 * 0.5},
        {"type_": "ImperviousSurface", "area": 1e7, "pollutant_load": DEPOSITION, "surface": "urban",
         "initial_storage": 5e6},
    ]
    nodes = [Waste(name="thames_above_abingdon")]
    nodes += [Catchment(name=c, data_input_dict=inputs[c]) for c in ("evenlode", "thames", "ray", "cherwell")]
    nodes += [
        ResidentialDemand(name="oxford", population=2e5, per_capita=per_capita, pollutant_load=household,
                          data_input_dict=inputs["oxford_land"]),
        Node(name="oxford_distribution"),
        Reservoir(name="farmoor", capacity=1e7, initial_storage=1e7, area=1.5e6, datum=62),
        FWTW(service_reservoir_storage_capacity=1e5, service_reservoir_storage_area=2e4,
             service_reservoir_initial_storage=0.9e5, treatment_throughput_capacity=4.5e4, name="oxford_fwtw"),
        WWTW(stormwater_storage_capacity=2e4, stormwater_storage_area=2e4, treatment_throughput_capacity=5e4,
             name="oxford_wwtw"),
        Sewer(capacity=4e6, pipe_timearea={0: 0.8, 1: 0.15, 2: 0.05}, name="combined_sewer"),
        Land(surfaces=surfaces, name="oxford_land", data_input_dict=inputs["oxford_land"]),
        Groundwater(capacity=3.2e9, area=3.2e8, name="gw", residence_time=20),
    ]
    nodes += [Node(name=n) for n in ("farmoor_abstraction", "evenlode_thames", "cherwell_ray", "cherwell_thames",
                                     "thames_mixer")]
    by_name = {n.name: n for n in nodes}
    # An arc registers itself with its two nodes when it is created, and that order is the order
    # in which a node later queries its arcs (nodes.py:257-265): create them in the demo's order ...
    spec = {name: (a, b, extra) for name, a, b, extra in ARCS}
    made = {name: Arc(in_port=by_name[spec[name][0]], out_port=by_name[spec[name][1]], name=name, **spec[name][2])
            for name in CREATION_ORDER}
    arcs = [made[name] for name, _, _, _ in ARCS]  # ... and list them in the order the model records them
    model = Model()
    model.add_instantiated_nodes(nodes)
    model.add_instantiated_arcs(arcs)
    model.dates = dates
    return model
