"""Flatten a WSIMOD object graph into the structure-of-arrays ``FlatModel``.

The flattener is duck-typed: it reads the attributes the reference classes define
(``wsimod/nodes/*.py``, ``wsimod/arcs/arcs.py``) and therefore works both on
``wsi_b200``'s own host-side mirror classes and on objects of the reference
package (that is how ``tests/golden/make_golden.py`` builds its fixtures).

Anything outside the supported hot path raises :class:`UnsupportedModelError` --
there is no CPU fallback (BASELINE.json north_star).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from . import layout as L


class UnsupportedModelError(RuntimeError):
    """The model uses a node/arc/orchestration feature outside the GPU hot path."""


SUPPORTED_NODES = set(L.NODE_TYPES)
SURFACE_CLASS_MAP = {
    "Surface": "Surface",
    "ImperviousSurface": "ImperviousSurface",
    "PerviousSurface": "PerviousSurface",
    "GrowingSurface": "GrowingSurface",
    # GardenSurface only differs through Demand push handlers (land.py:2384-2437);
    # Demand nodes are not on this path, so it behaves as a GrowingSurface.
    "GardenSurface": "GrowingSurface",
    # IrrigationSurface: a GrowingSurface that also pulls water through its Land (land.py:2314-2381);
    # VariableAreaSurface: a GrowingSurface whose climate scales with current_surface_area (land.py:2440-2461)
    "IrrigationSurface": "GrowingSurface",
    "VariableAreaSurface": "GrowingSurface",
}
STORE_TYPES = ("Groundwater", "Storage", "River", "Reservoir", "RiverReservoir")
ARC_CLASSES = ("Arc", "SewerArc", "WeirArc", "PullArc", "PushArc")  # arcs.py:18, 808-905 (no queues)

# (class name, orchestration function) -> op
ORCHESTRATION_OPS = {
    ("Land", "run"): "LAND_RUN",
    ("Land", "apply_irrigation"): "NOP",
    ("Groundwater", "infiltrate"): "GW_INFILTRATE",
    ("Groundwater", "distribute"): "GW_DISTRIBUTE",
    ("QueueGroundwater", "infiltrate"): "NOP",  # storage.py:365-367
    ("QueueGroundwater", "distribute"): "QGW_DISTRIBUTE",
    ("Sewer", "make_discharge"): "SEWER_DISCHARGE",
    ("River", "calculate_discharge"): "RIVER_CALC",
    ("River", "distribute"): "RIVER_DISTRIBUTE",
    ("Reservoir", "make_abstractions"): "RESERVOIR_ABSTRACT",
    ("Catchment", "route"): "CATCHMENT_ROUTE",
    ("Storage", "distribute"): "STORAGE_DISTRIBUTE",
    ("Reservoir", "distribute"): "STORAGE_DISTRIBUTE",
    ("RiverReservoir", "make_abstractions"): "RESERVOIR_ABSTRACT",
    ("RiverReservoir", "distribute"): "STORAGE_DISTRIBUTE",
    ("RiverReservoir", "satisfy_environmental"): "RESERVOIR_ENVIRONMENTAL",
    ("FWTW", "treat_water"): "FWTW_TREAT",
    ("Demand", "create_demand"): "DEMAND_CREATE",
    ("WWTW", "calculate_discharge"): "WWTW_CALC",
    ("WWTW", "make_discharge"): "WWTW_DISCHARGE",
}


@dataclass
class FlatModel:
    """Numpy image of a flattened model (see ``layout.py`` for the plane layout)."""

    pollutants: list
    additive: list
    non_additive: list
    float_accuracy: float
    R: int
    T: int
    nq: int
    # graph
    node_names: list
    node_type: np.ndarray
    node_gidx: np.ndarray
    arc_names: list
    arc_in: np.ndarray
    arc_out: np.ndarray
    out_ptr: np.ndarray
    out_arcs: np.ndarray
    in_ptr: np.ndarray
    in_arcs: np.ndarray
    cand_ptr: np.ndarray
    cand_arcs: np.ndarray
    land_surf_ptr: np.ndarray
    land_surf: np.ndarray
    surf_names: list  # (land name, surface name)
    sched_op: np.ndarray
    sched_node: np.ndarray
    # groups: name -> [n_fields, n_obj * R]
    groups: dict
    n_obj: dict
    # forcing
    node_forcing: np.ndarray  # [T, n_fseries, 3]
    surf_inputs: np.ndarray  # [n_months, n_dep_series, N_DEP]
    catch_inputs: np.ndarray  # [T, n_cseries, V]
    month_idx: np.ndarray
    dayofyear: np.ndarray
    doy_adj: np.ndarray
    arc_flags: np.ndarray = None  # [n_arcs] L.ARC_DENY_PUSH / L.ARC_DENY_PULL bits (PullArc / PushArc)
    dates: list = field(default_factory=list)
    recipe: dict = None  # what a synthetic builder drew (tests); not part of the model

    def __post_init__(self):
        if self.arc_flags is None:  # plain arcs
            self.arc_flags = np.zeros(max(len(self.arc_names), 1), dtype=np.int32)

    # ------------------------------------------------------------------ basics
    @property
    def P(self):
        return len(self.pollutants)

    @property
    def V(self):
        return self.P + 1

    @property
    def additive_mask(self):
        m = 0
        for p in self.additive:
            m |= 1 << (self.pollutants.index(p) + 1)
        return m

    @property
    def idx(self):
        return [
            (self.pollutants.index(n) + 1) if n in self.pollutants else -1 for n in L.NAMED
        ]

    def n_fields(self, group):
        if group == "sewer":
            return L.sewer_n_fields(self.V, self.nq)
        return L.GROUPS[group].n_fields(self.V)

    def plane(self, group, name, k=0):
        """Plane index of field ``name`` (+k) in ``group``."""
        return L.GROUPS[group].offset(name, self.V) + k

    def view(self, group, name, n=None):
        """View ``[n_planes, n_obj, R]`` of a field (writable)."""
        g = L.GROUPS[group]
        o = g.offset(name, self.V)
        if n is None:
            n = g.size(name, self.V)
        arr = self.groups[group]
        return arr[o : o + n].reshape(n, self.n_obj[group], self.R)

    # ------------------------------------------------------------- replication
    def replicate(self, R):
        """Return a copy with every object replicated ``R`` times (ensemble)."""
        if self.R != 1:
            raise ValueError("replicate() expects an R == 1 template")
        groups = {k: np.ascontiguousarray(np.repeat(v, R, axis=1)) for k, v in self.groups.items()}
        kw = dict(self.__dict__)
        kw.update(groups=groups, R=int(R))
        return FlatModel(**kw)

    def take_replicas(self, sel):
        """Sub-ensemble made of replicas ``sel`` (used to check a slice on the CPU)."""
        sel = np.asarray(sel)
        groups = {}
        for k, v in self.groups.items():
            n = self.n_obj[k]
            groups[k] = np.ascontiguousarray(
                v.reshape(v.shape[0], n, self.R)[:, :, sel].reshape(v.shape[0], n * len(sel))
            )
        kw = dict(self.__dict__)
        kw.update(groups=groups, R=int(len(sel)))
        return FlatModel(**kw)

    def copy(self):
        kw = dict(self.__dict__)
        kw["groups"] = {k: v.copy() for k, v in self.groups.items()}
        return FlatModel(**kw)

    # ------------------------------------------------------------------ probes
    def probes_for_arc(self, arc):
        a = self.arc_names.index(arc) if isinstance(arc, str) else int(arc)
        o = self.plane("arc", "vqip_in")
        return [(L.GROUP_ID["arc"], o + p, a) for p in range(self.V)]

    def probes_for_field(self, group, name, obj, n=None):
        g = L.GROUPS[group]
        if n is None:
            n = g.size(name, self.V)
        o = g.offset(name, self.V)
        return [(L.GROUP_ID[group], o + p, int(obj)) for p in range(n)]

    # --------------------------------------------------------------------- I/O
    _ARRAYS = [
        "node_type", "node_gidx", "arc_in", "arc_out", "out_ptr", "out_arcs", "in_ptr",
        "in_arcs", "cand_ptr", "cand_arcs", "land_surf_ptr", "land_surf", "sched_op",
        "sched_node", "node_forcing", "surf_inputs", "catch_inputs", "month_idx",
        "dayofyear", "doy_adj", "arc_flags",
    ]

    def to_npz_dict(self):
        d = {k: getattr(self, k) for k in self._ARRAYS}
        for k, v in self.groups.items():
            d["group_" + k] = v
        d["meta_pollutants"] = np.array(self.pollutants)
        d["meta_additive"] = np.array(self.additive)
        d["meta_non_additive"] = np.array(self.non_additive)
        d["meta_scalars"] = np.array([self.float_accuracy, self.R, self.T, self.nq], dtype=np.float64)
        d["meta_node_names"] = np.array(self.node_names)
        d["meta_arc_names"] = np.array(self.arc_names)
        d["meta_surf_names"] = np.array(["\t".join(x) for x in self.surf_names] or [""])
        d["meta_n_obj"] = np.array([self.n_obj[k] for k in L.GROUPS], dtype=np.int64)
        return d

    @classmethod
    def from_npz_dict(cls, d):
        sc = d["meta_scalars"]
        surf_names = [tuple(x.split("\t")) for x in d["meta_surf_names"].tolist() if x]
        kw = {k: np.ascontiguousarray(d[k]) for k in cls._ARRAYS if k in d}
        kw.setdefault("arc_flags", np.zeros(max(len(d["meta_arc_names"]), 1), dtype=np.int32))  # plain arcs
        V = len(d["meta_pollutants"]) + 1
        nq = int(sc[3])
        n_obj = {k: 0 for k in L.GROUPS}
        n_obj.update({k: int(v) for k, v in zip(L.GROUPS, d["meta_n_obj"])})
        groups = {}
        for k in L.GROUPS:
            nf = L.sewer_n_fields(V, nq) if k == "sewer" else L.GROUPS[k].n_fields(V)
            groups[k] = np.ascontiguousarray(d["group_" + k]) if "group_" + k in d else np.zeros((nf, 0))
        # fixtures written before a group gained planes are brought up to the current layout:
        if groups["surf"].shape[0] == L.SURF.n_fields(V) - 3 - V:  # irrigated_at: never, pre_irrigation: unused
            g = np.insert(groups["surf"], L.SURF.offset("irrigated_at", V), -1.0, axis=0)
            groups["surf"] = np.ascontiguousarray(np.concatenate([g, np.zeros((V, g.shape[1]))], axis=0))
        if groups["surf"].shape[0] == L.SURF.n_fields(V) - 2:  # irrigation_coefficient: none, climate_scale: 1
            o = L.SURF.offset("irrigation_coefficient", V)
            g = np.insert(groups["surf"], [o, o], 0.0, axis=0)
            g[o], g[o + 1] = -1.0, 1.0
            groups["surf"] = np.ascontiguousarray(g)
        if groups["tank"].shape[0] == L.TANK.n_fields(V) - 2:  # environmental_flow, env_satisfied: none
            o = L.TANK.offset("environmental_flow", V)
            groups["tank"] = np.ascontiguousarray(np.insert(groups["tank"], [o, o], 0.0, axis=0))
        if groups["distrib"].shape[0] == L.DISTRIB.n_fields(V) - 1:  # leakage: none
            groups["distrib"] = np.ascontiguousarray(np.insert(groups["distrib"], L.DISTRIB.offset("leakage", V), 0.0, axis=0))
        if groups["river"].shape[0] == L.RIVER.n_fields(V) - 1:  # mrf: none
            groups["river"] = np.ascontiguousarray(np.insert(groups["river"], L.RIVER.offset("mrf", V), 0.0, axis=0))
        if groups["land"].shape[0] == L.LAND.n_fields(V) - 1:  # garden_surf: none
            groups["land"] = np.ascontiguousarray(np.insert(groups["land"], L.LAND.offset("garden_surf", V), -1.0, axis=0))
        # the arc group's `stamp` (no arc carries anything) ...
        if groups["arc"].shape[0] == L.ARC.n_fields(V) - 1:
            groups["arc"] = np.ascontiguousarray(np.insert(groups["arc"], L.ARC.offset("stamp", V), -1.0, axis=0))
        # ... and the sewer group's pipe_timearea table (the default {0: 1}, sewer.py:62-66) and, later,
        # its QueueGroundwater / queue-decay planes (none)
        n_ta, n_q = 1 + 2 * L.MAX_TIMEAREA, 5 + 1 + 4 * L.MAX_DECAY
        full = L.sewer_n_fields(V, nq)
        if groups["sewer"].shape[0] == full - n_ta - n_q:
            o = L.SEWER.offset("n_timearea", V)
            ins = np.zeros((n_ta, groups["sewer"].shape[1]))
            ins[0] = 1.0
            ins[1 + L.MAX_TIMEAREA] = 1.0
            groups["sewer"] = np.concatenate([groups["sewer"][:o], ins, groups["sewer"][o:]], axis=0)
        if groups["sewer"].shape[0] == full - n_q:
            o = L.SEWER.offset("kind", V)
            ins = np.zeros((n_q, groups["sewer"].shape[1]))
            groups["sewer"] = np.ascontiguousarray(np.concatenate([groups["sewer"][:o], ins, groups["sewer"][o:]], axis=0))
        return cls(
            pollutants=[str(x) for x in d["meta_pollutants"].tolist()],
            additive=[str(x) for x in d["meta_additive"].tolist()],
            non_additive=[str(x) for x in d["meta_non_additive"].tolist()],
            float_accuracy=float(sc[0]),
            R=int(sc[1]),
            T=int(sc[2]),
            nq=int(sc[3]),
            node_names=[str(x) for x in d["meta_node_names"].tolist()],
            arc_names=[str(x) for x in d["meta_arc_names"].tolist()],
            surf_names=surf_names,
            groups=groups,
            n_obj=n_obj,
            **kw,
        )


# ---------------------------------------------------------------------- helpers
_HANDLER_DICTS = ("push_set_handler", "push_check_handler", "pull_set_handler", "pull_check_handler")


def _foreign(f):
    """A callable that is not code of the reference (or of this package's mirror of it)."""
    mod = getattr(getattr(f, "__func__", f), "__module__", None) or ""
    return not (mod.startswith("wsimod.") or mod.startswith("wsi_b200.") or mod == "builtins")


def _check_unpatched(obj, key, label):
    """SURVEY 8b eligibility: the kernels run the behaviour of the reference's classes, so an object whose
    behaviour was changed after construction must be refused, not silently run unpatched.  Caught here:
    ``wsimod.extensions.apply_patches`` (extensions.py:94-119: ``_patched_<method>`` attributes, replaced
    methods, replaced handler / list items), user functions assigned as methods or put into the handler
    dicts or a surface's inflows / processes / outflows lists, and handler dicts re-pointed at another
    reference method.  ``key`` indexes handler_table.HANDLER_TABLE (what a pristine object carries)."""
    if obj.__class__.__module__.startswith("wsi_b200."):
        return  # this package's own classes carry parameters and state only: there is no behaviour to patch
    import types

    from .handler_table import HANDLER_TABLE

    for k, v in vars(obj).items():
        if k.startswith("_patched_"):
            raise UnsupportedModelError(f"{label}: method {k[9:]!r} was patched (wsimod/extensions.py:115-119)")
        if isinstance(v, (types.MethodType, types.FunctionType)) and _foreign(v):
            raise UnsupportedModelError(f"{label}: attribute {k!r} is a user function; patched behaviour cannot run in the kernels")
        if isinstance(v, (list, dict)):
            for f in (v.values() if isinstance(v, dict) else v):
                if isinstance(f, (types.MethodType, types.FunctionType)) and _foreign(f):
                    raise UnsupportedModelError(f"{label}: {k!r} holds a user function; patched behaviour cannot run in the kernels")
    if key not in HANDLER_TABLE:
        return  # tanks, nutrient pools: the scan above is the whole check
    have = {h: {k: v.__name__ for k, v in getattr(obj, h).items()} for h in _HANDLER_DICTS if hasattr(obj, h)}
    aliases = {k: v.__name__ for k, v in vars(obj).items() if isinstance(v, types.MethodType)}
    ok = False
    for want in HANDLER_TABLE[key]:
        ok = ok or (all(have.get(h, want.get(h)) == want.get(h) for h in _HANDLER_DICTS) and aliases == want["aliases"])
    if not ok:
        raise UnsupportedModelError(f"{label}: handlers / method aliases differ from a pristine {key} ({have}, {aliases})")
    if hasattr(obj, "inflows"):
        names = {f.__name__ for lst in ("inflows", "processes", "outflows") for f in getattr(obj, lst)}
        extra = names - set(HANDLER_TABLE[key][0]["functions"])
        if extra:
            raise UnsupportedModelError(f"{label}: function lists hold {sorted(extra)}, not part of a pristine {key}")


def _cls(obj):
    name = obj.__class__.__name__
    # QueueGroundwater files itself under its parent's name (storage.py:299)
    if name == "Groundwater" and hasattr(obj, "timearea"):
        return "QueueGroundwater"
    # ... and so does RiverReservoir (storage.py:1041)
    if name == "Reservoir" and hasattr(obj, "environmental_flow"):
        return "RiverReservoir"
    return name


def _vq(vqip, pollutants):
    """dict VQIP -> list [volume, pollutants...]."""
    return [float(vqip.get("volume", 0.0))] + [float(vqip.get(p, 0.0)) for p in pollutants]


def _decay_fields(obj, pollutants):
    """(n, pol, const, exp, total, pre) for a DecayObj-like ``obj``."""
    decays = getattr(obj, "decays", None) or {}
    if len(decays) > L.MAX_DECAY:
        raise UnsupportedModelError(f"more than {L.MAX_DECAY} decaying pollutants on one store")
    n = len(decays)
    pol = [0.0] * L.MAX_DECAY
    const = [0.0] * L.MAX_DECAY
    exp_ = [0.0] * L.MAX_DECAY
    total = [0.0] * L.MAX_DECAY
    pre = [0.0] * L.MAX_DECAY
    td = getattr(obj, "total_decayed", None) or {}
    st_ = getattr(obj, "storage_", None) or {}
    for i, (p, pars) in enumerate(decays.items()):
        if p not in pollutants:
            raise UnsupportedModelError(f"decay for unknown pollutant {p}")
        pol[i] = float(pollutants.index(p) + 1)
        const[i] = float(pars["constant"])
        exp_[i] = float(pars["exponent"])
        total[i] = float(td.get(p, 0.0))
        pre[i] = float(st_.get(p, 0.0))
    return n, pol, const, exp_, total, pre


class _GroupBuilder:
    def __init__(self, name, V, nq=2):
        self.g = L.GROUPS[name]
        self.name = name
        self.V = V
        self.nf = L.sewer_n_fields(V, nq) if name == "sewer" else self.g.n_fields(V)
        self.cols = []

    def new(self):
        col = np.zeros(self.nf, dtype=np.float64)
        self.cols.append(col)
        return col

    def set(self, col, name, value, k=0):
        o = self.g.offset(name, self.V)
        if isinstance(value, (list, tuple, np.ndarray)):
            col[o + k : o + k + len(value)] = value
        else:
            col[o + k] = float(value)

    def array(self):
        if not self.cols:
            return np.zeros((self.nf, 0), dtype=np.float64)
        return np.ascontiguousarray(np.stack(self.cols, axis=1))


def _series_from_dict(did, dates, variables, required=True):
    out = np.zeros((len(dates), len(variables)), dtype=np.float64)
    for j, var in enumerate(variables):
        if var is None:
            continue
        for t, date in enumerate(dates):
            try:
                out[t, j] = did[(var, date)]
            except KeyError:
                if required:
                    raise UnsupportedModelError(f"missing input ({var!r}, {date})") from None
                out[t, j] = np.nan
    return out


PAD_PREFIX = "_pad"  # padding pollutants (see flatten): never shown to the caller


def model_pollutants(fm):
    """The model's own pollutants: fm.pollutants without the padding ones."""
    return [p for p in fm.pollutants if not p.startswith(PAD_PREFIX)]


# --------------------------------------------------------------------- flatten
def flatten(model, dates, pollutants, additive, non_additive, float_accuracy):
    """Flatten ``model`` for the timesteps ``dates`` (the objects used as
    ``data_input_dict`` keys, ``model.py:993-997``)."""
    pollutants = list(pollutants)
    additive = list(additive)
    non_additive = list(non_additive)
    P = len(pollutants)
    V = P + 1
    if P > L.MAXP:
        raise UnsupportedModelError(f"{P} pollutants > {L.MAXP}")
    if set(additive) | set(non_additive) != set(pollutants) or set(additive) & set(non_additive):
        raise UnsupportedModelError("POLLUTANTS must be the disjoint union of additive and non-additive")
    # The library is compiled for L.COMPILED_P pollutant counts.  Any other count runs on the next
    # compiled one with additive padding pollutants that are zero everywhere: a zero additive plane
    # stays zero through every VQIP operation (core.py:90-200) and no operation on another plane
    # reads it, so the model's own planes are bit-identical to an exact-width run.
    target = next((c for c in L.COMPILED_P if c >= P), None)
    if target is None:
        raise UnsupportedModelError(f"{P} pollutants > {max(L.COMPILED_P)}")
    pads = [f"{PAD_PREFIX}{i}" for i in range(target - P)]
    model_pollutants = list(pollutants)
    pollutants, additive = pollutants + pads, additive + pads
    P, V = len(pollutants), len(pollutants) + 1
    has_nitrate = "nitrate" in pollutants
    if has_nitrate:
        # River.biochemical_processes (storage.py:727) and the GrowingSurface nutrient processes read
        # the whole nutrient set once "nitrate" is modelled; other node classes never look at it
        needed = {"nitrite", "ammonia", "phosphate", "org-phosphorus", "org-nitrogen", "temperature"}
        uses_nutrients = any(
            _cls(n) == "River" or any(_cls(s) == "GrowingSurface" for s in getattr(n, "surfaces", []))
            for n in model.nodes.values()
        )
        if uses_nutrients and not needed <= set(pollutants):
            raise UnsupportedModelError(f"'nitrate' modelled without {sorted(needed - set(pollutants))}")
    dates = list(dates)
    T = len(dates)

    nodes = list(model.nodes.values())
    node_names = [n.name for n in nodes]
    node_index = {n.name: i for i, n in enumerate(nodes)}
    arcs = list(model.arcs.values())
    arc_names = [a.name for a in arcs]
    arc_index = {a.name: i for i, a in enumerate(arcs)}

    for n in nodes:
        if _cls(n) not in SUPPORTED_NODES:
            raise UnsupportedModelError(
                f"node {n.name!r} of class {_cls(n)} is outside the GPU hot path "
                f"(supported: {sorted(SUPPORTED_NODES)})"
            )
        _check_unpatched(n, _cls(n), f"node {n.name!r}")
        for s_ in getattr(n, "surfaces", []):
            _check_unpatched(s_, "surface:" + s_.__class__.__name__, f"surface {s_.surface!r} of {n.name!r}")
        for k_, v_ in vars(n).items():  # the node's tanks
            if hasattr(v_, "storage") and hasattr(v_, "capacity") and not isinstance(v_, (list, dict)):
                _check_unpatched(v_, "tank", f"{k_} of {n.name!r}")
    for a in arcs:
        if _cls(a) not in ARC_CLASSES:
            raise UnsupportedModelError(f"arc {a.name!r} of class {_cls(a)} is not supported (arcs without queues only)")
        _check_unpatched(a, "arc:" + _cls(a), f"arc {a.name!r}")
        irrigated = _cls(a.out_port) == "Land" and any(_cls(x) == "IrrigationSurface" for x in a.out_port.surfaces)
        if _cls(a.out_port) == "Catchment" or (_cls(a.out_port) == "Land" and _cls(a.in_port) != "Demand" and not irrigated):
            raise UnsupportedModelError(f"arc {a.name!r} pushes into a {_cls(a.out_port)} node")
        # Land.run only pushes to these types (land.py:200-244); an arc to a plain Storage carries nothing either way
        # (a Storage never pulls upstream); anything else downstream of a Land could pull through it -- refused
        if _cls(a.in_port) == "Land" and _cls(a.out_port) not in ("Sewer", "Groundwater", "QueueGroundwater", "River", "Node", "Waste", "Storage"):
            raise UnsupportedModelError(f"arc {a.name!r}: Land only pushes to Sewer / Groundwater / River / Node / Waste")

    # ---- graph CSR -----------------------------------------------------------
    node_type = np.array([L.NT[_cls(n)] for n in nodes], dtype=np.int32)
    arc_in = np.array([node_index[a.in_port.name] for a in arcs], dtype=np.int32)
    arc_out = np.array([node_index[a.out_port.name] for a in arcs], dtype=np.int32)

    def csr(lists):
        ptr = np.zeros(len(lists) + 1, dtype=np.int32)
        flat = []
        for i, l in enumerate(lists):
            flat.extend(l)
            ptr[i + 1] = len(flat)
        return ptr, np.array(flat if flat else [0], dtype=np.int32)[: max(len(flat), 1)]

    # PullArc answers every push (check) with "nothing" and PushArc every pull (arcs.py:808-893): an arc whose
    # check returns zero never takes part in a distribution (nodes.py:271-317), so a PullArc is simply not a
    # push candidate; pulls read arc_flags on the device
    arc_flags = np.array([(L.ARC_DENY_PUSH if _cls(a) == "PullArc" else 0) | (L.ARC_DENY_PULL if _cls(a) == "PushArc" else 0)
                          for a in arcs] or [0], dtype=np.int32)
    pushable = lambda k: not (arc_flags[arc_index[k]] & L.ARC_DENY_PUSH)  # noqa: E731
    out_lists = [[arc_index[k] for k in n.out_arcs.keys()] for n in nodes]
    in_lists = [[arc_index[k] for k in n.in_arcs.keys()] for n in nodes]
    out_ptr, out_arcs = csr(out_lists)
    in_ptr, in_arcs = csr(in_lists)
    cand_lists = []
    for n in nodes:
        for fname in L.FILTER_NAMES:
            types = L.FILTERS[fname]
            if types is None:
                cand_lists.append([arc_index[k] for k in n.out_arcs.keys() if pushable(k)])
            else:
                # get_direction_arcs (nodes.py:257-265): per type, in of_type order
                lst = []
                for ty in types:
                    lst += [arc_index[k] for k in n.out_arcs_type.get(ty, {}).keys() if pushable(k)]
                cand_lists.append(lst)
    for n_i, n in enumerate(nodes):
        longest = max(len(l) for l in cand_lists[n_i * L.NF : (n_i + 1) * L.NF])
        if longest > L.MAX_CAND:
            # the kernels split a push over at most MAX_CAND arcs (nodes.py:440-473 has no limit)
            raise UnsupportedModelError(
                f"node {n.name!r} has {longest} eligible out-arcs for one push; the kernels support {L.MAX_CAND}"
            )
    cand_ptr, cand_arcs = csr(cand_lists)

    # ---- forcing series ------------------------------------------------------
    fseries_ids = {}
    fseries = []

    def forcing_series(did, variables):
        key = (id(did), tuple(variables))
        if key not in fseries_ids:
            fseries_ids[key] = len(fseries)
            fseries.append(_series_from_dict(did, dates, variables))
        return fseries_ids[key]

    months = []
    month_of = {}
    month_idx = np.zeros(T, dtype=np.int32)
    for t, date in enumerate(dates):
        m = date.to_period("M")
        if m not in month_of:
            month_of[m] = len(months)
            months.append(m)
        month_idx[t] = month_of[m]
    dep_ids = {}
    dep_series = []

    def dep_series_id(did, need):
        key = id(did)
        if key not in dep_ids:
            arr = np.zeros((len(months), L.N_DEP), dtype=np.float64)
            present = set(x[0] for x in did.keys())
            for j, var in enumerate(L.DEP_COLS):
                if var in present:
                    for mi, m in enumerate(months):
                        try:
                            arr[mi, j] = did[(var, m)]
                        except KeyError:
                            raise UnsupportedModelError(f"surface input ({var!r}, {m}) missing") from None
                elif var in need:
                    raise UnsupportedModelError(f"surface data_input_dict lacks {var!r}")
            dep_ids[key] = len(dep_series)
            dep_series.append(arr)
        return dep_ids[key]

    dayofyear = np.array([int(d.dayofyear) for d in dates], dtype=np.int32)
    # land.py:1524-1529: `if self.parent.t.is_leap_year:` -- a bound method (csv/yaml
    # dates, model.py:114) is always truthy, a pandas bool is not.
    doy_adj = dayofyear.copy()
    for t, d in enumerate(dates):
        if bool(d.is_leap_year) and doy_adj[t] > 59:
            doy_adj[t] -= 1

    # ---- groups --------------------------------------------------------------
    sewers = [n for n in nodes if _cls(n) == "Sewer"]
    nq = 2
    for w in sewers:
        keys = list(w.sewer_tank.internal_arc.queue.keys())
        ta = getattr(w, "pipe_timearea", None) or {0: 1}
        nq = max(nq, max([1, int(w.pipe_time)] + [int(k) for k in keys] + [int(k) for k in ta]) + 1)
    for w in nodes:
        if _cls(w) == "QueueGroundwater":
            delay = int(w.tank.internal_arc.number_of_timesteps)
            keys = [int(k) for k in w.tank.internal_arc.queue.keys()]
            nq = max(nq, max([1] + keys + [int(k) + delay for k in w.timearea]) + 1)
    if nq > 16:
        raise UnsupportedModelError("sewer travel time > 15 timesteps")

    gb = {k: _GroupBuilder(k, V, nq) for k in L.GROUPS}
    node_gidx = np.zeros(len(nodes), dtype=np.int32)
    land_surf_lists = []
    surf_names = []
    catch_series = []

    for a in arcs:
        col = gb["arc"].new()
        gb["arc"].set(col, "capacity", a.capacity)
        gb["arc"].set(col, "preference", a.preference)
        gb["arc"].set(col, "flow_in", a.flow_in)
        gb["arc"].set(col, "vqip_in", _vq(a.vqip_in, pollutants))
        carried = a.flow_in != 0 or any(x != 0 for x in _vq(a.vqip_in, pollutants))
        gb["arc"].set(col, "stamp", 0.0 if carried else -1.0)  # an arc flattened mid-step keeps what it carries

    def add_store(n):
        col = gb["tank"].new()
        tk = n.tank
        b = gb["tank"]
        b.set(col, "capacity", tk.capacity)
        b.set(col, "area", tk.area)
        b.set(col, "residence_time", getattr(n, "residence_time", 1.0))
        b.set(col, "infiltration_threshold", getattr(n, "infiltration_threshold", 1.0))
        b.set(col, "infiltration_pct", getattr(n, "infiltration_pct", 0.0))
        is_decay = _cls(tk) == "DecayTank"
        if _cls(tk) not in ("Tank", "DecayTank"):
            raise UnsupportedModelError(f"{n.name}: tank class {_cls(tk)} unsupported")
        b.set(col, "is_decay", 1.0 if is_decay else 0.0)
        b.set(col, "environmental_flow", getattr(n, "environmental_flow", 0.0))
        b.set(col, "env_satisfied", getattr(n, "total_environmental_satisfied", 0.0))
        need_T = (is_decay and len(tk.decays) > 0) or (_cls(n) == "River" and has_nitrate)
        fid = 0
        if need_T:
            fid = forcing_series(n.data_input_dict, (None, None, "temperature"))
        b.set(col, "forcing_id", fid)
        if is_decay:
            nd, pol, const, exp_, total, pre = _decay_fields(tk, pollutants)
            b.set(col, "n_decay", nd)
            b.set(col, "decay_pol", pol)
            b.set(col, "decay_const", const)
            b.set(col, "decay_exp", exp_)
            b.set(col, "total_decayed", total)
            b.set(col, "pre_decay", pre)
        b.set(col, "storage", _vq(tk.storage, pollutants))
        b.set(col, "storage_", _vq(tk.storage_, pollutants))
        return len(gb["tank"].cols) - 1

    for i, n in enumerate(nodes):
        c = _cls(n)
        if c == "Land":
            node_gidx[i] = len(gb["land"].cols)
            col = gb["land"].new()
            b = gb["land"]
            b.set(col, "res_surface", n.surface_runoff.residence_time)
            b.set(col, "res_subsurface", n.subsurface_runoff.residence_time)
            b.set(col, "res_percolation", n.percolation.residence_time)
            b.set(col, "forcing_id", forcing_series(n.data_input_dict, ("precipitation", "et0", "temperature")))
            b.set(col, "node", i)
            b.set(col, "garden_surf", -1.0)
            b.set(col, "surface_runoff", _vq(n.surface_runoff.storage, pollutants))
            b.set(col, "subsurface_runoff", _vq(n.subsurface_runoff.storage, pollutants))
            b.set(col, "percolation", _vq(n.percolation.storage, pollutants))
            lst = []
            for s in n.surfaces:
                lst.append(len(gb["surf"].cols))
                surf_names.append((n.name, str(s.surface)))
                _flatten_surface(gb["surf"], s, pollutants, additive, has_nitrate, dep_series_id)
                gb["surf"].set(gb["surf"].cols[-1], "land", node_gidx[i])
                if _cls(s) == "GardenSurface":  # the LAST garden answers the ("Demand", "Garden") tag (land.py:120-128)
                    b.set(col, "garden_surf", float(lst[-1]))
            land_surf_lists.append(lst)
        elif c == "QueueGroundwater":
            node_gidx[i] = len(gb["sewer"].cols)
            _flatten_queue_groundwater(gb["sewer"], n, pollutants, V, forcing_series)
        elif c in STORE_TYPES:
            ti = add_store(n)
            if c == "River":
                node_gidx[i] = len(gb["river"].cols)
                _flatten_river(gb["river"], n, ti)
            else:
                node_gidx[i] = ti
        elif c == "Sewer":
            node_gidx[i] = len(gb["sewer"].cols)
            col = gb["sewer"].new()
            b = gb["sewer"]
            tk = n.sewer_tank
            if any(_cls(a.out_port) == "Land" for a in n.out_arcs.values()):
                raise UnsupportedModelError(f"{n.name}: sewer flooding onto Land (sewer.py:236) is unsupported")
            if int(n.pipe_time) != n.pipe_time or n.pipe_time < 0:
                raise UnsupportedModelError("non-integer pipe_time")
            b.set(col, "capacity", tk.capacity)
            b.set(col, "pipe_time", n.pipe_time)
            ta = getattr(n, "pipe_timearea", None) or {0: 1}
            if len(ta) > L.MAX_TIMEAREA or any(int(k) != k or k < 0 for k in ta):
                raise UnsupportedModelError(f"{n.name}: pipe_timearea with more than {L.MAX_TIMEAREA} / non-integer entries")
            b.set(col, "n_timearea", len(ta))
            b.set(col, "ta_time", [float(k) for k in ta.keys()])
            b.set(col, "ta_frac", [float(v) for v in ta.values()])
            qmask = 0
            for k, v in tk.internal_arc.queue.items():
                qmask |= 1 << int(k)
                b.set(col, "queue", _vq(v, pollutants), k=int(k) * V)
            b.set(col, "qmask", qmask)
            b.set(col, "storage", _vq(tk.storage, pollutants))
            b.set(col, "storage_", _vq(tk.storage_, pollutants))
            b.set(col, "active_storage", _vq(tk.active_storage, pollutants))
        elif c == "Catchment":
            node_gidx[i] = len(gb["catchm"].cols)
            col = gb["catchm"].new()
            col[0] = len(catch_series)
            catch_series.append(_series_from_dict(n.data_input_dict, dates, ["flow"] + model_pollutants + [None] * len(pads)))
        elif c in ("WWTW", "FWTW"):
            node_gidx[i] = len(gb["works"].cols)
            _flatten_works(gb["works"], n, add_store, pollutants, additive, non_additive)
        elif c == "Demand":
            residential = hasattr(n, "population")
            if not residential and n.__class__.__mro__[0].__name__ == "NonResidentialDemand":
                raise UnsupportedModelError(f"{n.name}: NonResidentialDemand.get_demand calls itself (demand.py:143-150)")
            missing = [p for p in model_pollutants if p not in n.pollutant_load]
            if residential and missing:  # the reference raises KeyError in get_house_demand (demand.py:303-324)
                raise UnsupportedModelError(f"{n.name}: pollutant_load lacks {missing}")
            unknown = [p for p in n.pollutant_load if p not in model_pollutants]
            if unknown:  # Demand.get_demand would put a foreign key into the VQIP (demand.py:126-129)
                raise UnsupportedModelError(f"{n.name}: pollutant_load names {unknown}, not modelled")
            node_gidx[i] = len(gb["demand"].cols)
            b = gb["demand"]
            col = b.new()
            if residential:
                for f in ("population", "per_capita", "gardening_efficiency", "constant_weighting", "constant_temp"):
                    b.set(col, f, getattr(n, f))
                b.set(col, "forcing_id", forcing_series(n.data_input_dict, (None, None, "temperature")))
            else:  # base Demand (demand.py:14-131): population -1 marks it, per_capita holds constant_demand,
                b.set(col, "population", -1.0)  # pollutant_load the absolute loads of that demand
                b.set(col, "per_capita", float(n.constant_demand))
            b.set(col, "pollutant_load", [0.0] + [float(n.pollutant_load.get(p, 0.0)) for p in pollutants])
            for f in ("total_demand", "total_backup", "total_received"):
                b.set(col, f, _vq(getattr(n, f), pollutants))
        elif c == "UnlimitedDistribution":
            node_gidx[i] = len(gb["distrib"].cols)
            col = gb["distrib"].new()
            gb["distrib"].set(col, "supplied", _vq(n.supplied, pollutants))
        elif c == "Distribution":
            if not 0 <= n.leakage < 1:
                raise UnsupportedModelError(f"{n.name}: leakage must be in [0, 1)")
            node_gidx[i] = len(gb["distrib"].cols)
            col = gb["distrib"].new()
            gb["distrib"].set(col, "leakage", float(n.leakage))
        else:  # Node, Waste
            node_gidx[i] = 0

    land_surf_ptr, land_surf = csr(land_surf_lists)

    # ---- schedule ------------------------------------------------------------
    sched = []
    for item in model.orchestration:
        for type_key, fn in item.items():
            for n in model.nodes_type.get(type_key, {}).values():
                op = ORCHESTRATION_OPS.get((_cls(n), fn))
                if op == "NOP" and fn == "apply_irrigation" and any(_cls(x) == "IrrigationSurface" for x in n.surfaces):
                    op = "LAND_IRRIGATE"
                if op is None:
                    raise UnsupportedModelError(f"orchestration entry {type_key}.{fn} on {_cls(n)} {n.name!r} unsupported")
                if op != "NOP":
                    sched.append((L.OP[op], node_index[n.name]))
    for name in model.river_discharge_order:
        sched.append((L.OP["RIVER_DISTRIBUTE"], node_index[name]))
    sched_op = np.array([s[0] for s in sched] or [L.OP["NOP"]], dtype=np.int32)
    sched_node = np.array([s[1] for s in sched] or [0], dtype=np.int32)

    groups = {k: b.array() for k, b in gb.items()}
    n_obj = {k: v.shape[1] for k, v in groups.items()}
    node_forcing = (
        np.stack(fseries, axis=1) if fseries else np.zeros((T, 1, 3), dtype=np.float64)
    )
    node_forcing = np.nan_to_num(np.ascontiguousarray(node_forcing), nan=0.0)
    surf_inputs = (
        np.stack(dep_series, axis=1) if dep_series else np.zeros((max(len(months), 1), 1, L.N_DEP))
    )
    catch_inputs = (
        np.stack(catch_series, axis=1) if catch_series else np.zeros((T, 1, V), dtype=np.float64)
    )
    return FlatModel(
        pollutants=pollutants,
        additive=additive,
        non_additive=non_additive,
        float_accuracy=float(float_accuracy),
        R=1,
        T=T,
        nq=nq,
        node_names=node_names,
        node_type=node_type,
        node_gidx=node_gidx,
        arc_names=arc_names,
        arc_in=arc_in,
        arc_out=arc_out,
        out_ptr=out_ptr,
        out_arcs=out_arcs,
        in_ptr=in_ptr,
        in_arcs=in_arcs,
        cand_ptr=cand_ptr,
        cand_arcs=cand_arcs,
        land_surf_ptr=land_surf_ptr,
        land_surf=land_surf,
        surf_names=surf_names,
        arc_flags=arc_flags,
        sched_op=sched_op,
        sched_node=sched_node,
        groups=groups,
        n_obj=n_obj,
        node_forcing=np.ascontiguousarray(node_forcing, dtype=np.float64),
        surf_inputs=np.ascontiguousarray(surf_inputs, dtype=np.float64),
        catch_inputs=np.ascontiguousarray(catch_inputs, dtype=np.float64),
        month_idx=month_idx,
        dayofyear=dayofyear,
        doy_adj=doy_adj,
        dates=dates,
    )


def _flatten_queue_groundwater(b, n, pollutants, V, forcing_series):
    """QueueGroundwater (storage.py:248-449): a QueueTank -- or, with decays, a DecayQueueTank whose
    internal DecayArcAlt decays what enters and what waits (tanks.py:645-696, arcs.py:731-805) -- in
    the queue-store group."""
    tk = n.tank
    ia = tk.internal_arc
    ta = n.timearea or {0: 1}
    if len(ta) > L.MAX_TIMEAREA or any(int(k) != k or k < 0 for k in ta):
        raise UnsupportedModelError(f"{n.name}: timearea with more than {L.MAX_TIMEAREA} / non-integer entries")
    if len(n.in_arcs) and any(_cls(a.in_port) in ("Reservoir", "RiverReservoir", "FWTW", "Demand") for a in n.out_arcs.values()):
        raise UnsupportedModelError(f"{n.name}: abstraction from a QueueGroundwater (pull_set_active) is not supported")
    # push checks are answered by the Tank the Storage base class created first (storage.py:60): stale
    handler = getattr(n, "push_check_handler", {}).get("default")
    stale = getattr(handler, "__self__", None) or getattr(n, "_stale_tank", None) or tk
    col = b.new()
    b.set(col, "capacity", tk.capacity)
    b.set(col, "pipe_time", 0)
    b.set(col, "n_timearea", len(ta))
    b.set(col, "ta_time", [float(k) for k in ta.keys()])
    b.set(col, "ta_frac", [float(v) for v in ta.values()])
    b.set(col, "kind", 1)
    b.set(col, "queue_delay", int(ia.number_of_timesteps))
    b.set(col, "stale_capacity", stale.capacity)
    b.set(col, "stale_volume", stale.storage["volume"])
    decays = getattr(ia, "decays", None) or {}
    if decays:
        nd, pol, const, exp_, total, _ = _decay_fields(ia, pollutants)
        b.set(col, "n_decay", nd)
        b.set(col, "decay_pol", pol)
        b.set(col, "decay_const", const)
        b.set(col, "decay_exp", exp_)
        b.set(col, "total_decayed", total)
        b.set(col, "forcing_id", forcing_series(n.data_input_dict, (None, None, "temperature")))
    qmask = 0
    for k, v in ia.queue.items():
        qmask |= 1 << int(k)
        b.set(col, "queue", _vq(v, pollutants), k=int(k) * V)
    b.set(col, "qmask", qmask)
    b.set(col, "storage", _vq(tk.storage, pollutants))
    b.set(col, "storage_", _vq(tk.storage_, pollutants))
    b.set(col, "active_storage", _vq(tk.active_storage, pollutants))


def _flatten_works(b, n, add_store, pollutants, additive, non_additive):
    """WWTW / FWTW (wtw.py:14-657): process parameters, state and the works' own tank."""
    if "temperature" not in pollutants:
        raise UnsupportedModelError(f"{n.name}: treatment works read the influent temperature (wtw.py:138)")
    for key in additive:
        if key not in n.process_parameters or key not in n.liquor_multiplier:
            raise UnsupportedModelError(f"{n.name}: process_parameters / liquor_multiplier lack {key!r}")
    kind = 1 if _cls(n) == "FWTW" else 0

    class _TankHolder:  # the works' tank as a Storage-like object for add_store
        pass

    h = _TankHolder()
    h.tank = n.service_reservoir_tank if kind else n.stormwater_tank
    h.residence_time = h.infiltration_threshold = h.infiltration_pct = 0.0
    h.data_input_dict = None
    h.name = n.name
    col = b.new()
    b.set(col, "kind", kind)
    b.set(col, "tank", add_store(h))
    b.set(col, "throughput_capacity", n.treatment_throughput_capacity)
    b.set(col, "percent_solids", n.percent_solids)
    pidx = {p: i + 1 for i, p in enumerate(pollutants)}
    lm = [0.0] * (len(pollutants) + 1)
    pc = [0.0] * (len(pollutants) + 1)
    pe = [1.0] * (len(pollutants) + 1)
    lm[0] = float(n.liquor_multiplier["volume"])
    pc[0] = float(n.process_parameters["volume"]["constant"])
    for key in additive:
        lm[pidx[key]] = float(n.liquor_multiplier[key])
        pc[pidx[key]] = float(n.process_parameters[key]["constant"])
        pe[pidx[key]] = float(n.process_parameters[key]["exponent"])
    b.set(col, "liquor_mult", lm)
    b.set(col, "proc_const", pc)
    b.set(col, "proc_exp", pe)
    for f in ("current_input", "treated", "liquor", "solids"):
        b.set(col, f, _vq(getattr(n, f), pollutants))
    if kind:
        for f in ("total_deficit", "total_pulled", "previous_pulled", "unpushed_sludge"):
            b.set(col, f, _vq(getattr(n, f), pollutants))
    else:
        b.set(col, "liquor_", _vq(n.liquor_, pollutants))


def _flatten_surface(b, s, pollutants, additive, has_nitrate, dep_series_id):
    cname = _cls(s)
    if cname not in SURFACE_CLASS_MAP:
        raise UnsupportedModelError(
            f"surface class {cname} is outside the GPU hot path (pull-driven irrigation / variable area)"
        )
    kind = SURFACE_CLASS_MAP[cname]
    col = b.new()
    b.set(col, "type", L.ST[kind])
    b.set(col, "irrigated_at", -1.0)
    b.set(col, "irrigation_coefficient", float(s.irrigation_coefficient) if cname == "IrrigationSurface" else -1.0)
    b.set(col, "climate_scale", s.current_surface_area / s.area if cname == "VariableAreaSurface" else 1.0)
    b.set(col, "area", s.area)
    b.set(col, "capacity", s.capacity)
    b.set(col, "depth", s.depth)
    load = getattr(s, "pollutant_load", {}) or {}
    b.set(col, "has_load", 1.0 if len(load) > 0 else 0.0)
    lv = [0.0] + [float(load.get(p, 0.0)) for p in pollutants]
    b.set(col, "load", lv)
    did = getattr(s, "data_input_dict", {}) or {}
    present = set(x[0] for x in did.keys())
    has_dep = "nhx-dry" in present
    b.set(col, "has_dep", 1.0 if has_dep else 0.0)
    need = set()
    if has_dep:
        need |= set(L.DEP_COLS[:6])
    if kind == "GrowingSurface" and has_nitrate:
        need |= set(L.DEP_COLS[6:])
    if need:
        b.set(col, "dep_id", dep_series_id(did, need))
    nd, pol, const, exp_, total, pre = _decay_fields(s, pollutants)
    b.set(col, "n_decay", nd)
    b.set(col, "decay_pol", pol)
    b.set(col, "decay_const", const)
    b.set(col, "decay_exp", exp_)
    b.set(col, "total_decayed", total)
    b.set(col, "pre_decay", pre)
    b.set(col, "storage", _vq(s.storage, pollutants))
    if kind == "ImperviousSurface":
        b.set(col, "et0_to_e", s.et0_to_e)
    if kind in ("PerviousSurface", "GrowingSurface"):
        for name in (
            "field_capacity_m", "wilting_point_m", "infiltration_capacity", "surface_coefficient",
            "percolation_coefficient", "subsurface_coefficient", "ihacres_p", "et0_coefficient",
            "soil_temp_w_prev", "soil_temp_w_air", "soil_temp_w_deep", "soil_temp_deep",
        ):
            b.set(col, name, getattr(s, name))
    if kind == "GrowingSurface":
        for name in (
            "rooting_depth", "ET_depletion_factor", "total_available_water", "readily_available_water",
            "sowing_day", "harvest_day", "satact", "thetaupp", "thetalow", "thetapow", "uptake1",
            "uptake2", "uptake3", "uptake_PNratio", "erodibility", "sreroexp", "cohesion", "slope",
            "srfilt", "macrofilt", "limpar", "exppar", "hsatINs", "denpar", "kfr", "nfr", "kadsdes",
            "bulk_density", "crop_cover", "ground_cover", "crop_factor",
        ):
            b.set(col, name, getattr(s, name))
        b.set(col, "adsorption_nr_limit", s.adosorption_nr_limit)  # sic (land.py:1243)
        b.set(col, "adsorption_nr_maxiter", s.adsorption_nr_maxiter)
        b.set(col, "autumn_sow", 1.0 if s.autumn_sow else 0.0)
        ns = len(s.crop_factor_stages)
        if ns > L.MAX_STAGE or ns != len(s.crop_factor_stage_dates):
            raise UnsupportedModelError("crop_factor_stages table too long / inconsistent")
        b.set(col, "n_stage", ns)
        b.set(col, "stage_dates", [float(x) for x in s.crop_factor_stage_dates])
        b.set(col, "stage_vals", [float(x) for x in s.crop_factor_stages])
        b.set(col, "hs_calendar", [float(x) for x in s.harvest_sow_calendar])
        b.set(col, "ground_cover_stages", [float(x) for x in s.ground_cover_stages])
        b.set(col, "crop_cover_stages", [float(x) for x in s.crop_cover_stages])
        b.set(col, "days_after_sow", -1.0 if s.days_after_sow is None else float(s.days_after_sow))
        npool = s.nutrient_pool
        b.set(col, "frac_dry_n_to_dis_inorg", npool.fraction_dry_n_to_dissolved_inorganic)
        b.set(col, "frac_dry_n_to_fast", npool.fraction_dry_n_to_fast)
        NP = ("N", "P")
        b.set(col, "frac_manure_to_dis_inorg", [npool.fraction_manure_to_dissolved_inorganic[x] for x in NP])
        b.set(col, "frac_manure_to_fast", [npool.fraction_manure_to_fast[x] for x in NP])
        for name in ("degrhpar", "dishpar", "minfpar", "disfpar", "immobdpar"):
            b.set(col, name, [getattr(npool, name)[x] for x in NP])
        b.set(col, "pool_fast", [npool.fast_pool.storage[x] for x in NP])
        b.set(col, "pool_humus", [npool.humus_pool.storage[x] for x in NP])
        b.set(col, "pool_dis_inorg", [npool.dissolved_inorganic_pool.storage[x] for x in NP])
        b.set(col, "pool_dis_org", [npool.dissolved_organic_pool.storage[x] for x in NP])
        b.set(col, "pool_ads_inorg", [npool.adsorbed_inorganic_pool.storage[x] for x in NP])


def _flatten_river(b, n, tank_index):
    col = b.new()
    b.set(col, "tank", tank_index)
    # get_riverrc (storage.py:900-915), evaluated with the host libm so the
    # constant is bit-identical to the reference's
    total_time = n.length / n.velocity
    kt = n.damp * total_time
    riverrc = 1 - kt + kt * math.exp(-1 / kt) if kt != 0 else 1
    b.set(col, "riverrc", riverrc)
    b.set(col, "mrf", getattr(n, "mrf", 0.0))
    b.set(col, "area", n.area)
    for name in (
        "uptake_PNratio", "denpar_w", "T_wdays", "halfsatINwater", "hsatTP", "limpppar", "prodNpar",
        "prodPpar", "muptNpar", "muptPpar", "max_temp_lag", "max_phosphorus_lag",
    ):
        b.set(col, name, getattr(n, name))
    if n.max_temp_lag + 1 > L.LAG_T or n.max_phosphorus_lag + 1 > L.LAG_TP:
        raise UnsupportedModelError("river lag longer than the ring buffers")
    if getattr(n, "mrf", 0) not in (0, 0.0):
        pass  # mrf only limits pulls (storage.py:676-684), which are not on this path
    lt = list(n.lagged_temperatures)
    ltp = list(n.lagged_total_phosphorus)
    b.set(col, "n_lag_T", len(lt))
    b.set(col, "head_T", 0)
    b.set(col, "n_lag_TP", len(ltp))
    b.set(col, "head_TP", 0)
    if lt:
        b.set(col, "lag_T", [float(x) for x in lt])
    if ltp:
        b.set(col, "lag_TP", [float(x) for x in ltp])
    b.set(col, "bio_in", [0.0])
    b.set(col, "bio_out", [0.0])


# ------------------------------------------------------------------ scatter back
def scatter_back(model, fm, groups, last_date=None):
    """Write the final SoA state (``groups``: name -> ``[n_fields, n_obj]`` arrays of an
    ``R == 1`` run) back into the Python objects, so that follow-on ``run()`` calls,
    ``node.tank.storage[...]`` inspection and pickling keep working."""
    if fm.R != 1:
        raise ValueError("scatter_back needs an R == 1 model")
    pol = model_pollutants(fm)  # the padding pollutants never reach the objects (they lead nowhere: all zero)
    V = fm.V

    def vq_dict(group, name, obj, k=0):
        o = L.GROUPS[group].offset(name, V) + k * V
        col = groups[group][o : o + V, obj]
        d = {"volume": float(col[0])}
        d.update({p: float(col[i + 1]) for i, p in enumerate(pol)})
        return d

    def sc(group, name, obj, k=0):
        return float(groups[group][L.GROUPS[group].offset(name, V) + k, obj])

    def decay_back(tank, group, obj):
        decays = getattr(tank, "decays", None) or {}
        tank.total_decayed = dict.fromkeys(pol + ["volume"], 0)
        for i, p in enumerate(decays):
            tank.total_decayed[p] = sc(group, "total_decayed", obj, i)
            tank.storage_[p] = sc(group, "pre_decay", obj, i)

    si = ti = ri = wi = li = 0
    for n in model.nodes.values():
        c = _cls(n)
        n.t = last_date
        if c == "Land":
            for name in ("surface_runoff", "subsurface_runoff", "percolation"):
                tank = getattr(n, name)
                tank.storage = vq_dict("land", name, li)
                tank.storage_ = dict(tank.storage)
            for s in n.surfaces:
                s.storage = vq_dict("surf", "storage", si)
                s.storage_ = dict(s.storage)
                decay_back(s, "surf", si)
                kind = SURFACE_CLASS_MAP[_cls(s)]
                if kind in ("PerviousSurface", "GrowingSurface"):
                    s.tank_recharge = sc("surf", "out_tank_recharge", si)
                    for attr, f in (("percolation", "out_percolation"), ("subsurface_flow", "out_subsurface"),
                                    ("infiltration_excess", "out_surface"), ("evaporation", "out_evaporation"),
                                    ("precipitation", "out_precipitation")):
                        d = dict.fromkeys(pol + ["volume"], 0)
                        d["volume"] = sc("surf", f, si)
                        setattr(s, attr, d)
                if kind == "GrowingSurface":
                    das = sc("surf", "days_after_sow", si)
                    s.days_after_sow = None if das < 0 else int(das)
                    s.crop_cover = sc("surf", "crop_cover", si)
                    s.ground_cover = sc("surf", "ground_cover", si)
                    s.crop_factor = sc("surf", "crop_factor", si)
                    s.et0_coefficient = sc("surf", "et0_coefficient", si)
                    npool = s.nutrient_pool
                    for pool, f in ((npool.fast_pool, "pool_fast"), (npool.humus_pool, "pool_humus"),
                                    (npool.dissolved_inorganic_pool, "pool_dis_inorg"),
                                    (npool.dissolved_organic_pool, "pool_dis_org"),
                                    (npool.adsorbed_inorganic_pool, "pool_ads_inorg")):
                        pool.storage["N"] = sc("surf", f, si, 0)
                        pool.storage["P"] = sc("surf", f, si, 1)
                si += 1
            li += 1
        elif c in STORE_TYPES:
            n.tank.storage = vq_dict("tank", "storage", ti)
            n.tank.storage_ = vq_dict("tank", "storage_", ti)
            if _cls(n.tank) == "DecayTank":
                decay_back(n.tank, "tank", ti)
            if c == "RiverReservoir":
                n.total_environmental_satisfied = sc("tank", "env_satisfied", ti)
            if c == "River":
                nT, hT = int(sc("river", "n_lag_T", ri)), int(sc("river", "head_T", ri))
                n.lagged_temperatures = [sc("river", "lag_T", ri, (hT + k) % L.LAG_T) for k in range(nT)]
                nP, hP = int(sc("river", "n_lag_TP", ri)), int(sc("river", "head_TP", ri))
                n.lagged_total_phosphorus = [sc("river", "lag_TP", ri, (hP + k) % L.LAG_TP) for k in range(nP)]
                ri += 1
            ti += 1
        elif c == "Sewer":
            tk = n.sewer_tank
            tk.storage = vq_dict("sewer", "storage", wi)
            tk.storage_ = vq_dict("sewer", "storage_", wi)
            tk.active_storage = vq_dict("sewer", "active_storage", wi)
            qmask = int(sc("sewer", "qmask", wi))
            tk.internal_arc.queue = {
                k: vq_dict("sewer", "queue", wi, k) for k in range(fm.nq) if qmask & (1 << k)
            }
            tk.internal_arc.max_travel = max([1] + list(tk.internal_arc.queue))
            wi += 1
    for a in model.arcs.values():
        a.flow_in = 0
        a.flow_out = 0
        a.vqip_in = dict.fromkeys(pol + ["volume"], 0)
        a.vqip_out = a.vqip_in
