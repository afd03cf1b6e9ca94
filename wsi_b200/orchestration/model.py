"""``Model``: the drop-in for ``wsimod.orchestration.model.Model`` on the land/tank/arc path.

``Model.load`` / ``add_nodes`` / ``add_arcs`` build the same object graph from the same YAML
files and dicts (reference: wsimod/orchestration/model.py:216-274, 642-778).  ``Model.run``
(reference :893-1200) does not loop over Python objects: it flattens the graph into fp64
structure-of-arrays tensors, steps them with the CUDA kernels through the C ABI, and returns
the same ``(flows, tanks, objective_results, surfaces)`` records.  Models that use features
outside the supported path are refused (``UnsupportedModelError``) -- there is no CPU
fallback.
"""
from __future__ import annotations

import csv
import gzip
import os
from collections.abc import Sequence
from datetime import datetime

import numpy as np
import yaml

from ..arcs import arcs as arcs_mod
from ..backend import flatten as F
from ..backend import layout as L
from ..backend.flatten import UnsupportedModelError
from ..backend.probes import SURF_OUT, ProbeSet
from ..core import constants
from ..core.core import WSIObj
from ..nodes.nodes import NODES_REGISTRY
from ..nodes.tanks import QueueTank, ResidenceTank, Tank


class to_datetime:
    """Date wrapper used for csv / YAML dates (reference: model.py:40-170).  ``is_leap_year``
    is a *method* there, so ``if t.is_leap_year:`` is always true (land.py:1526); the
    flattener evaluates that truthiness on the live date objects, which keeps both date
    regimes reproducible."""

    def __init__(self, date_string):
        self._date = self._parse(date_string)
        self._monthly = len(date_string) == 7 and date_string.count("-") == 1

    @staticmethod
    def _parse(s):
        for fmt in ("%Y-%m-%d %H:%M:%S", "%Y-%m-%d", "%Y-%m"):
            try:
                return datetime.strptime(s, fmt)
            except ValueError:
                continue
        raise ValueError(f"Unable to parse date string '{s}'. Expected 'YYYY-MM-DD', 'YYYY-MM' or 'YYYY-MM-DD HH:MM:SS'")

    def __str__(self):
        return self._date.strftime("%Y-%m" if self._monthly else "%Y-%m-%d")

    __repr__ = __str__

    @property
    def dayofyear(self):
        return self._date.timetuple().tm_yday

    @property
    def day(self):
        return self._date.day

    @property
    def year(self):
        return self._date.year

    @property
    def month(self):
        return self._date.month

    def to_period(self, args="M"):
        return to_datetime(f"{self._date.year}-{str(self._date.month).zfill(2)}")

    def is_leap_year(self):
        y = self._date.year
        return y % 4 == 0 and (y % 100 != 0 or y % 400 == 0)

    def __eq__(self, other):
        return isinstance(other, to_datetime) and self._date == other._date

    def __hash__(self):
        return hash(self._date)


def open_func(file_path, mode):
    if mode == "rt" and file_path.endswith(".gz"):
        return gzip.open(file_path, mode)
    return open(file_path, mode)


def read_csv(file_path, delimiter=","):
    """``{(variable, to_datetime(time)): value}`` from a long-format csv(.gz) (model.py:1250-1268)."""
    with open_func(file_path, "rt") as f:
        return {(row["variable"], to_datetime(row["time"])): float(row["value"])
                for row in csv.DictReader(f, delimiter=delimiter)}


class _Records(Sequence):
    """Lazily materialised list of result dicts (one per recorded element per timestep)."""

    def __init__(self, n_steps, makers):
        self._n_steps = n_steps
        self._makers = makers  # callables (t) -> dict, in within-step order

    def __len__(self):
        return self._n_steps * len(self._makers)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        n = len(self)
        if i < 0:
            i += n
        if not 0 <= i < n:
            raise IndexError(i)
        t, k = divmod(i, len(self._makers))
        return self._makers[k](t)

    def __iter__(self):
        for t in range(self._n_steps):
            for mk in self._makers:
                yield mk(t)

    def __eq__(self, other):
        return list(self) == list(other)


class Model(WSIObj):
    def __init__(self):
        super().__init__()
        self.arcs = {}
        self.nodes = {}
        self.nodes_type = {}
        self.extensions = []
        self.river_discharge_order = []
        self.nodelist = []
        # default orchestration (model.py:191-205)
        self.orchestration = [
            {"FWTW": "treat_water"},
            {"Demand": "create_demand"},
            {"Land": "run"},
            {"Groundwater": "infiltrate"},
            {"Sewer": "make_discharge"},
            {"Foul": "make_discharge"},
            {"WWTW": "calculate_discharge"},
            {"Groundwater": "distribute"},
            {"River": "calculate_discharge"},
            {"Reservoir": "make_abstractions"},
            {"Land": "apply_irrigation"},
            {"WWTW": "make_discharge"},
            {"Catchment": "route"},
        ]
        self.last_run = None

    # ------------------------------------------------------------------ building
    def load(self, address, config_name="config.yml", overrides=None):
        """Load a YAML model file and its csv inputs (model.py:216-274)."""
        with open(os.path.join(address, config_name), "r") as file:
            data = yaml.safe_load(file)
        for key, item in (overrides or {}).items():
            data[key] = item
        constants.POLLUTANTS = data.get("pollutants", constants.POLLUTANTS)
        constants.ADDITIVE_POLLUTANTS = data.get("additive_pollutants", constants.ADDITIVE_POLLUTANTS)
        constants.NON_ADDITIVE_POLLUTANTS = data.get("non_additive_pollutants", constants.NON_ADDITIVE_POLLUTANTS)
        constants.FLOAT_ACCURACY = float(data.get("float_accuracy", constants.FLOAT_ACCURACY))
        self.__dict__.update(Model().__dict__)
        if data.get("extensions"):
            # user patches replace node methods (wsimod/extensions.py:65-119); patched behaviour
            # cannot run inside the kernels, so refuse instead of ignoring it
            raise UnsupportedModelError("YAML 'extensions' (node patches) are not supported on the GPU path")
        if "orchestration" in data:
            self.orchestration = data["orchestration"]
        if "nodes" not in data:
            raise ValueError("No nodes found in the config")
        nodes = data["nodes"]
        if data.get("unified_data_file"):
            self._load_unified_data(address, data["unified_data_file"], nodes, data)
        else:
            self._load_individual_files(address, nodes, data)
        self.add_nodes(list(nodes.values()))
        self.add_arcs(list(data.get("arcs", {}).values()))
        self.add_overrides(data.get("overrides", {}))

    def _load_individual_files(self, address, nodes, data):
        for node in nodes.values():
            if "filename" in node:
                node["data_input_dict"] = read_csv(os.path.join(address, node.pop("filename")))
            if "surfaces" in node:
                for surface in node["surfaces"].values():
                    if "filename" in surface:
                        surface["data_input_dict"] = read_csv(os.path.join(address, surface.pop("filename")))
                node["surfaces"] = list(node["surfaces"].values())
        if "dates" in data:
            self.dates = [to_datetime(x) for x in data["dates"]]

    def _load_unified_data(self, address, unified_data_file, nodes, data):
        """Single parquet file with columns node/surface/variable/time/value (model.py:276-326);
        keys become pandas Timestamps / monthly Periods."""
        import pandas as pd

        df = pd.read_parquet(os.path.join(address, unified_data_file))
        surf = df.dropna(subset=["surface"]).copy()
        surf["time"] = pd.to_datetime(surf["time"]).dt.to_period("M")
        node_df = df.loc[df.surface.isna()].copy()
        node_df["time"] = pd.to_datetime(node_df["time"])
        surface_dict = {k: g.set_index(["variable", "time"]).value.to_dict() for k, g in surf.groupby(["node", "surface"])}
        node_dict = {k: g.set_index(["variable", "time"]).value.to_dict() for k, g in node_df.groupby("node")}
        for name, node in nodes.items():
            if node.get("data_input_dict"):
                node["data_input_dict"] = node_dict[name]
            if "surfaces" in node:
                for key, surface in node["surfaces"].items():
                    if surface.get("data_input_dict"):
                        surface["data_input_dict"] = surface_dict[(name, key)]
                node["surfaces"] = list(node["surfaces"].values())
        if "dates" in data:
            self.dates = list(pd.to_datetime(data["dates"]))

    def add_nodes(self, nodelist):
        """Instantiate nodes from parameter dicts (model.py:642-672)."""
        for data in nodelist:
            data = dict(data)
            name = data["name"]
            type_ = data.pop("type_")
            node_type = data.pop("node_type_override", type_)
            if "foul" in name:
                type_ = "Foul"  # the reference files foul sewers under their own key (model.py:658)
            data.pop("geometry", None)
            if node_type not in NODES_REGISTRY:
                raise UnsupportedModelError(
                    f"Node type {node_type} is not available on the GPU path (have {sorted(NODES_REGISTRY)})"
                )
            self.nodes_type.setdefault(type_, {})[name] = NODES_REGISTRY[node_type](**data)
            self.nodes[name] = self.nodes_type[type_][name]
        self.nodelist = list(self.nodes.values())

    def add_instantiated_nodes(self, nodelist):
        self.nodelist = list(nodelist)
        self.nodes = {x.name: x for x in nodelist}
        for x in nodelist:
            self.nodes_type.setdefault(x.__class__.__name__, {})[x.name] = x

    def add_arcs(self, arclist):
        """Instantiate arcs from parameter dicts and derive the river order (model.py:690-735)."""
        for arc in arclist:
            arc = dict(arc)
            type_ = arc.pop("type_")
            if type_ not in arcs_mod.SUPPORTED_ARCS:
                raise UnsupportedModelError(f"Arc type {type_} is not available on the GPU path (arcs without queues only)")
            arc["in_port"] = self.nodes[arc["in_port"]]
            arc["out_port"] = self.nodes[arc["out_port"]]
            self.arcs[arc["name"]] = getattr(arcs_mod, type_)(**arc)
        self._derive_river_order()

    def add_instantiated_arcs(self, arclist):
        self.arclist = list(arclist)
        self.arcs = {x.name: x for x in arclist}
        self._derive_river_order()

    _RIVER_CLASSES = ("River", "Node", "Waste", "Reservoir")

    def _derive_river_order(self):
        """Upstream-first order of River nodes: breadth-first levels from the Waste nodes over
        arcs between River/Node/Waste/Reservoir nodes, most upstream level first.  The
        reference does this with a recursive set-based sweep (model.py:720-735, 780-807) that
        is hash-ordered within a level and recursion-limited; this is the iterative
        equivalent with a deterministic within-level order."""
        self.river_discharge_order = []
        river_arcs = [a for a in self.arcs.values()
                      if a.in_port.__class__.__name__ in self._RIVER_CLASSES
                      and a.out_port.__class__.__name__ in self._RIVER_CLASSES]
        if not river_arcs:
            return
        level = {n: 0 for n in self.nodes_type.get("Waste", {})}
        upstream_of = {}
        for a in river_arcs:
            upstream_of.setdefault(a.out_port.name, []).append(a.in_port.name)
        frontier = list(level)
        lvl = 0
        while frontier:
            lvl += 1
            nxt = []
            for n in frontier:
                for up in upstream_of.get(n, []):
                    if up not in level:
                        level[up] = lvl
                        nxt.append(up)
            frontier = nxt
        rivers = self.nodes_type.get("River", {})
        order = sorted((n for n in level if n in rivers), key=lambda n: -level[n])
        self.river_discharge_order = order

    def add_overrides(self, config):
        for node in config.get("nodes", {}).values():
            node = dict(node)
            type_, name = node.pop("type_"), node.pop("name")
            if type_ not in self.nodes_type:
                raise ValueError(f"Node type {type_} not recognised")
            if name not in self.nodes_type[type_]:
                raise ValueError(f"Node {name} not recognised")
            self.nodes_type[type_][name].apply_overrides(node)
        for arc in config.get("arcs", {}).values():
            arc = dict(arc)
            name = arc.pop("name")
            arc.pop("type_", None)
            if name not in self.arcs:
                raise ValueError(f"Arc {name} not recognised")
            self.arcs[name].apply_overrides(arc)

    def default_settings(self):
        return {"arcs": {"flows": True, "pollutants": True}, "tanks": {"storages": True, "pollutants": True},
                "mass_balance": False}

    # ---------------------------------------------------------------------- run
    def flatten(self, dates=None):
        """SoA image of the current object state for ``dates`` (default: ``self.dates``)."""
        dates = self.dates if dates is None else dates
        return F.flatten(self, list(dates), constants.POLLUTANTS, constants.ADDITIVE_POLLUTANTS,
                         constants.NON_ADDITIVE_POLLUTANTS, constants.FLOAT_ACCURACY)

    def run(self, dates=None, settings=None, record_arcs=None, record_tanks=None, record_surfaces=None,
            verbose=True, record_all=True, objectives=[], device="cuda:0"):
        """Simulate ``dates`` on the GPU (reference semantics: model.py:893-1200).

        Returns ``(flows, tanks, objective_results, surfaces)``; the three record collections
        behave as lists of dicts with the reference's keys and are materialised lazily."""
        from ..backend.engine import GpuRunner

        if record_arcs is None:
            record_arcs = list(self.arcs.keys()) if record_all else []
        record_arcs = list(record_arcs)
        record_tanks = list(record_tanks or [])
        record_surfaces = list(record_surfaces or [])
        dates = list(self.dates if dates is None else dates)
        for objective in objectives:
            kind = objective["element_type"]
            if kind == "tanks":
                record_tanks.append(objective["name"])
            elif kind == "flows":
                record_arcs.append(objective["name"])
            elif kind == "surfaces":
                record_surfaces.append((objective["name"], objective["surface"]))
            else:
                print("element_type not recorded")

        fm = self.flatten(dates)
        ps = ProbeSet(fm)
        ps.add_flows(record_arcs)
        plan = _RecordPlan(self, fm, ps, record_tanks, record_surfaces, record_all)
        runner = GpuRunner(fm, probes=ps.probes, rec_t0=0, rec_T=len(dates), record_mb=False, device=device)
        runner.run(0, len(dates))
        runner.synchronize()
        out = ps.split(runner.out_probe())
        flags = runner.node_flags()
        self.last_run = {"node_flags": flags, "mb_count": runner.mb_count(), "mb_sums": runner.mb_sums(),
                         "launches": runner.launch_count()}
        if verbose:
            self._print_advisories(fm, flags, runner.mb_count())
        F.scatter_back(self, fm, runner.groups(), dates[-1] if dates else None)

        flows, tanks, surfaces = plan.records(out, dates)
        objective_results = []
        for objective in objectives:
            kind = objective["element_type"]
            if kind == "tanks":
                sel = [x for x in tanks if x["node"] == objective["name"]]
            elif kind == "flows":
                sel = [x for x in flows if x["arc"] == objective["name"]]
            else:
                sel = [x for x in surfaces if x["node"] == objective["name"] and x["surface"] == objective["surface"]]
            objective_results.append(objective["function"](sel, self))
        return flows, tanks, objective_results, surfaces

    def _print_advisories(self, fm, flags, mb_count):
        """The reference prints per-object advisories (core.py:423-427, nodes.py:476); the
        kernels reduce them to sticky per-node status bits."""
        from ..backend.capi import FLAG_NAMES

        per_node = np.bitwise_or.reduce(flags.reshape(len(fm.node_names), fm.R), axis=1)
        for name, f in zip(fm.node_names, per_node):
            if f:
                what = ", ".join(v for k, v in FLAG_NAMES.items() if f & k)
                print(f"advisory: {name}: {what}")
        n = int(np.count_nonzero(mb_count))
        if n:
            print(f"mass balance tolerance exceeded on {n} timestep(s)")

    def reinit(self):
        for node in self.nodes.values():
            node.reinit()

    # ---- persistence (model.py:604-640) and result files (__main__.py:59-63) -----------------
    def save_pickle(self, fid):
        """Pickle the whole object graph including its state (``run`` scatters the device state
        back into the objects, so a model saved after a run resumes where it stopped)."""
        import dill as pickle

        with open(fid, "wb") as file:
            pickle.dump(self, file)

    def load_pickle(self, fid):
        import dill as pickle

        with open(fid, "rb") as file:
            return pickle.load(file)

    @staticmethod
    def write_results(flows, tanks, surfaces, outputs):
        """``flows.csv`` / ``tanks.csv`` / ``surfaces.csv`` exactly as the reference's command line
        writes them (``pd.DataFrame(records).to_csv``, __main__.py:59-63)."""
        import os

        import pandas as pd

        os.makedirs(outputs, exist_ok=True)
        for name, records in (("flows", flows), ("tanks", tanks), ("surfaces", surfaces)):
            pd.DataFrame(list(records)).to_csv(os.path.join(str(outputs), name + ".csv"))


class _RecordPlan:
    """Which probes back ``tanks`` / ``surfaces`` records and how to turn them into the
    reference's dicts (model.py:1063-1170)."""

    def __init__(self, model, fm, ps, record_tanks, record_surfaces, record_all):
        self.model, self.fm, self.ps = model, fm, ps
        self.pol = F.model_pollutants(fm)  # results never carry the padding pollutants
        V = fm.V
        self.tank_items = []  # (kind, node name, prop or None, block name, index)
        self.surf_items = []
        store_index = {}
        ti = 0
        for n in model.nodes.values():
            if n.__class__.__name__ in F.STORE_TYPES:
                store_index[n.name] = ti
                ti += 1
        sewer_index = {n.name: i for i, n in enumerate(x for x in model.nodes.values() if x.__class__.__name__ == "Sewer")}
        land_index = {n.name: i for i, n in enumerate(x for x in model.nodes.values() if x.__class__.__name__ == "Land")}
        surf_index = {key: i for i, key in enumerate(fm.surf_names)}
        want_tank, want_sewer, want_land, want_surf = [], [], [], []

        def tank_ref(node, prop_name, prop):
            if node.__class__.__name__ == "Land":
                k = ("surface_runoff", "subsurface_runoff", "percolation").index(prop_name)
                want_land.append(land_index[node.name])
                return ("land", land_index[node.name], k)
            if node.__class__.__name__ == "Sewer":
                want_sewer.append(sewer_index[node.name])
                return ("sewer", sewer_index[node.name], 0)
            want_tank.append(store_index[node.name])
            return ("tank", store_index[node.name], 0)

        for name in record_tanks:
            node = model.nodes[name]
            self.tank_items.append(("brief", name, None, tank_ref(node, "tank", node.tank)))
        if record_all:
            for node in model.nodes.values():
                for prop_ in dir(node):
                    prop = getattr(node, prop_, None)
                    if prop.__class__ in (QueueTank, Tank, ResidenceTank):
                        self.tank_items.append(("full", node.name, prop_, tank_ref(node, prop_, prop)))
        surf_list = list(record_surfaces)
        if record_all:
            surf_list += [(key[0], key[1]) for key in fm.surf_names]
        for node_name, surface_name in surf_list:
            i = surf_index[(node_name, str(surface_name))]
            want_surf.append(i)
            self.surf_items.append((node_name, surface_name, i))
        self.sel = {
            "tank": sorted(set(want_tank)), "sewer": sorted(set(want_sewer)), "land": sorted(set(want_land)),
            "surf": sorted(set(want_surf)),
        }
        if self.sel["tank"]:
            ps.add_field("rec_tank", "tank", "storage", self.sel["tank"])
        if self.sel["sewer"]:
            ps.add_field("rec_sewer", "sewer", "storage", self.sel["sewer"])
        if self.sel["land"]:
            ps.add_fields("rec_land", "land", ["surface_runoff", "subsurface_runoff", "percolation"], self.sel["land"])
        if self.sel["surf"]:
            ps.add_fields("rec_surf", "surf", ["storage"] + SURF_OUT + ["capacity", "et0_coefficient"], self.sel["surf"])
        self.V = V

    def records(self, out, dates):
        fm, V, pol = self.fm, self.V, self.pol
        n_steps = len(dates)
        flows_arr = out.get("flows")
        flow_arcs = [fm.arc_names[a] for a in self.ps.flow_arcs]

        def flow_maker(k, name):
            def mk(t):
                v = flows_arr[t, k, :, 0]
                d = {"arc": name, "flow": float(v[0]), "time": dates[t]}
                d.update({p: float(v[i + 1]) for i, p in enumerate(pol)})
                return d
            return mk

        flows = _Records(n_steps, [flow_maker(k, n) for k, n in enumerate(flow_arcs)])

        def vq_of(ref, t):
            grp, idx, k = ref
            arr = out["rec_" + grp]
            j = self.sel[grp].index(idx)
            return arr[t, j, k * V:(k + 1) * V, 0]

        def tank_maker(kind, name, prop, ref):
            def mk(t):
                v = vq_of(ref, t)
                if kind == "brief":
                    return {"node": name, "storage": float(v[0]), "time": dates[t]}
                d = {"node": name, "time": dates[t], "storage": float(v[0]), "prop": prop}
                d.update({p: float(v[i + 1]) for i, p in enumerate(pol)})
                return d
            return mk

        tanks = _Records(n_steps, [tank_maker(*it) for it in self.tank_items])
        types = fm.view("surf", "type")[0, :, 0]

        def surf_maker(node_name, surface_name, i):
            j = self.sel["surf"].index(i)
            impervious = int(types[i]) == L.ST["ImperviousSurface"]

            def mk(t):
                row = out["rec_surf"][t, j, :, 0]
                st, o = row[:V], row[V:]
                if impervious:
                    d = {"node": node_name, "surface": surface_name, "storage": float(st[0]),
                         "evaporation": float(o[3]), "precipitation": float(o[4]), "capacity": float(o[6]),
                         "time": dates[t]}
                else:
                    d = {"node": node_name, "surface": surface_name, "percolation": float(o[0]),
                         "subsurface_r": float(o[1]), "surface_r": float(o[2]), "storage": float(st[0]),
                         "evaporation": float(o[3]), "precipitation": float(o[4]), "tank_recharge": float(o[5]),
                         "capacity": float(o[6]), "time": dates[t], "et0_coef": float(o[7])}
                d.update({p: float(st[k + 1]) for k, p in enumerate(pol)})
                return d
            return mk

        surfaces = _Records(n_steps, [surf_maker(*it) for it in self.surf_items])
        return flows, tanks, surfaces


__all__ = ["Model", "to_datetime", "read_csv", "UnsupportedModelError"]
