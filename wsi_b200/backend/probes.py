"""Named recording sets: which planes to copy out of the state every timestep.

``Model.run`` records arc flows, tank storages and surface diagnostics before
``end_timestep`` (``wsimod/orchestration/model.py:1063-1170``).  On the device
this is one mechanism: a list of probes ``(group, plane, object)`` copied into
``out_probe[t, k, r]``.  ``ProbeSet`` builds such lists and turns the recorded
block back into named arrays.
"""
from __future__ import annotations

import numpy as np

from . import layout as L

SURF_OUT = [
    "out_percolation", "out_subsurface", "out_surface", "out_evaporation", "out_precipitation",
    "out_tank_recharge",
]
SURF_STATE = [
    "et0_coefficient", "days_after_sow", "crop_cover", "ground_cover", "crop_factor",
]
SURF_POOLS = ["pool_fast", "pool_humus", "pool_dis_inorg", "pool_dis_org", "pool_ads_inorg"]


class ProbeSet:
    def __init__(self, fm):
        self.fm = fm
        self.probes = []
        self.blocks = {}  # name -> (start, shape)

    def _add(self, name, probes, shape):
        self.blocks[name] = (len(self.probes), tuple(shape))
        self.probes.extend(probes)

    def add_flows(self, arcs=None):
        """VQIP on every (or the listed) arc: ``flows[t, arc, V]`` (model.py:1063-1069)."""
        fm = self.fm
        arcs = list(range(len(fm.arc_names))) if arcs is None else [
            fm.arc_names.index(a) if isinstance(a, str) else int(a) for a in arcs
        ]
        pr = []
        for a in arcs:
            pr += fm.probes_for_arc(a)
        self._add("flows", pr, (len(arcs), fm.V))
        self.flow_arcs = arcs
        return self

    def add_field(self, name, group, field, objs=None, n=None):
        fm = self.fm
        objs = list(range(fm.n_obj[group])) if objs is None else list(objs)
        if n is None:
            n = L.GROUPS[group].size(field, fm.V)
        pr = []
        for o in objs:
            pr += fm.probes_for_field(group, field, o, n)
        self._add(name, pr, (len(objs), n))
        return self

    def add_fields(self, name, group, fields, objs=None):
        fm = self.fm
        objs = list(range(fm.n_obj[group])) if objs is None else list(objs)
        pr = []
        width = 0
        for o in objs:
            width = 0
            for f in fields:
                n = L.GROUPS[group].size(f, fm.V)
                pr += fm.probes_for_field(group, f, o, n)
                width += n
        self._add(name, pr, (len(objs), width))
        return self

    def add_standard(self):
        """Everything the parity tests compare."""
        fm = self.fm
        self.add_flows()
        if fm.n_obj["land"]:
            self.add_fields("land_tanks", "land", ["surface_runoff", "subsurface_runoff", "percolation"])
        if fm.n_obj["surf"]:
            self.add_field("surf_storage", "surf", "storage")
            self.add_fields("surf_out", "surf", SURF_OUT)
            self.add_fields("surf_state", "surf", SURF_STATE + SURF_POOLS)
        # the storage of every store node in node order (what Model.run's tank record walks): the tank of
        # a Groundwater / Storage / Reservoir / River, the queue store of a QueueGroundwater
        store_probes, n_stores = [], 0
        for i, t in enumerate(fm.node_type):
            name = L.NODE_TYPES[int(t)]
            g = int(fm.node_gidx[i])
            if name in ("Groundwater", "Storage", "Reservoir", "RiverReservoir"):
                store_probes += fm.probes_for_field("tank", "storage", g)
            elif name == "River":
                store_probes += fm.probes_for_field("tank", "storage", int(fm.view("river", "tank")[0, g, 0]))
            elif name == "QueueGroundwater":
                store_probes += fm.probes_for_field("sewer", "storage", g)
            else:
                continue
            n_stores += 1
        if n_stores:
            self._add("tank_storage", store_probes, (n_stores, fm.V))
        sewers = [int(fm.node_gidx[i]) for i, t in enumerate(fm.node_type) if L.NODE_TYPES[int(t)] == "Sewer"]
        if sewers:
            self.add_field("sewer_storage", "sewer", "storage", objs=sewers)
            self.add_field("sewer_active", "sewer", "active_storage", objs=sewers)
        return self

    def split(self, out_probe):
        """``out_probe[T, n_probe, R]`` -> dict of ``[T, n_obj, width, R]`` arrays."""
        out = {}
        out_probe = np.asarray(out_probe)
        T, _, R = out_probe.shape
        for name, (start, shape) in self.blocks.items():
            n = int(np.prod(shape))
            out[name] = out_probe[:, start : start + n, :].reshape((T,) + shape + (R,))
        return out
