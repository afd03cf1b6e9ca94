"""ctypes binding of ``libwsi_b200.so`` (include/wsi_b200.h).

``WsiModelDesc`` mirrors the C struct field for field.  ``ModelBuffers`` owns every
buffer a descriptor points at -- torch CUDA tensors for the product path, numpy
arrays when the CPU oracle (tests only) is handed the same descriptor -- and keeps
them alive for as long as the descriptor is in use.

There is no CPU fallback here: :func:`load_library` raises if the CUDA extension
has not been built, and ``ModelBuffers(device="cuda")`` raises without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import layout as L

ABI_VERSION = 2

RUN_SCHEDULE = 1
RUN_MASS_BALANCE = 2
RUN_RECORD = 4
RUN_END_TIMESTEP = 8
RUN_SURFACES = 16
RUN_ALL = 31

FLAG_NAMES = {
    1: "mass_balance",
    2: "maxiter",
    4: "pool_mismatch",
    8: "nan",
    16: "junction_depth",
    32: "push_denied",
    64: "supply_deficit",
}

_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)
_u32p = C.POINTER(C.c_uint32)


class WsiModelDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("P", C.c_int32),
        ("V", C.c_int32),
        ("R", C.c_int32),
        ("T", C.c_int32),
        ("n_nodes", C.c_int32),
        ("n_arcs", C.c_int32),
        ("n_land", C.c_int32),
        ("n_surf", C.c_int32),
        ("n_tank", C.c_int32),
        ("n_river", C.c_int32),
        ("n_sewer", C.c_int32),
        ("n_catch", C.c_int32),
        ("n_sched", C.c_int32),
        ("nq", C.c_int32),
        ("n_fseries", C.c_int32),
        ("n_dep_series", C.c_int32),
        ("n_months", C.c_int32),
        ("n_cseries", C.c_int32),
        ("n_probe", C.c_int32),
        ("rec_t0", C.c_int32),
        ("rec_T", C.c_int32),
        ("record_mb", C.c_int32),
        ("additive_mask", C.c_uint32),
        ("idx", C.c_int32 * 8),
        ("float_accuracy", C.c_double),
        ("node_type", _i32p),
        ("node_gidx", _i32p),
        ("arc_in", _i32p),
        ("arc_out", _i32p),
        ("out_ptr", _i32p),
        ("out_arcs", _i32p),
        ("in_ptr", _i32p),
        ("in_arcs", _i32p),
        ("cand_ptr", _i32p),
        ("cand_arcs", _i32p),
        ("land_surf_ptr", _i32p),
        ("land_surf", _i32p),
        ("surf_order", _i32p),
        ("surf_class_ptr", C.c_int32 * 5),
        ("features", C.c_int32),
        ("end_kind", _i32p),
        ("end_obj", _i32p),
        ("n_end", C.c_int32),
        ("reserved1", C.c_int32),
        ("sched_op", _i32p),
        ("sched_node", _i32p),
        ("engine", C.c_int32),
        ("n_net_levels", C.c_int32),
        ("n_net_calc", C.c_int32),
        ("n_net_final", C.c_int32),
        ("n_net_bnd", C.c_int32),
        ("reserved2", C.c_int32),
        ("net_level_ptr", _i32p),
        ("net_level_ptr_host", _i32p),
        ("net_level_node", _i32p),
        ("net_calc_node", _i32p),
        ("net_final_node", _i32p),
        ("net_bnd_node", _i32p),
        ("net_path_ptr", _i32p),
        ("net_path_split", _i32p),
        ("net_path_sender", _i32p),
        ("net_path_arc_ptr", _i32p),
        ("net_path_arcs", _i32p),
        ("net_item", _i32p),
        ("net_scratch", _i32p),
        ("net_done", _i32p),
        ("net_rpath", _i32p),
        ("arc", _f64p),
        ("land", _f64p),
        ("surf", _f64p),
        ("tank", _f64p),
        ("river", _f64p),
        ("sewer", _f64p),
        ("catchm", _f64p),
        ("node_forcing", _f64p),
        ("surf_inputs", _f64p),
        ("catch_inputs", _f64p),
        ("month_idx", _i32p),
        ("dayofyear", _i32p),
        ("doy_adj", _i32p),
        ("probe_group", _i32p),
        ("probe_plane", _i32p),
        ("probe_obj", _i32p),
        ("out_probe", _f64p),
        ("out_mb", _f64p),
        ("mb_terms", _f64p),
        ("mb_term_slot", _i32p),
        ("mb_partial", _f64p),
        ("mb_sums", _f64p),
        ("mb_count", _i32p),
        ("node_flags", _u32p),
        ("arc_chain_ptr", _i32p),
        ("arc_chain", _i32p),
        ("n_red", C.c_int32),
        ("rec_r0", C.c_int32),
        ("rec_R", C.c_int32),
        ("reserved3", C.c_int32),
        ("red_group", _i32p),
        ("red_plane", _i32p),
        ("red_obj", _i32p),
        ("red_out", _f64p),
        ("n_works", C.c_int32),
        ("n_demand", C.c_int32),
        ("works", _f64p),
        ("demand", _f64p),
        ("distrib", _f64p),
        ("n_distrib", C.c_int32),
        ("sched_split", C.c_int32),
        ("arc_flags", _i32p),
    ]


_NET_ARRAYS = ["net_level_ptr", "net_level_node", "net_calc_node", "net_final_node", "net_bnd_node", "net_path_ptr",
               "net_path_split", "net_path_sender", "net_path_arc_ptr", "net_path_arcs", "net_item", "net_scratch", "net_done", "net_rpath"]
_INT_ARRAYS = _NET_ARRAYS + [
    "arc_chain_ptr", "arc_chain", "surf_order", "end_kind", "end_obj", "mb_term_slot", "node_type", "node_gidx", "arc_in", "arc_out", "out_ptr", "out_arcs", "in_ptr", "in_arcs",
    "cand_ptr", "cand_arcs", "land_surf_ptr", "land_surf", "sched_op", "sched_node", "month_idx",
    "dayofyear", "doy_adj", "arc_flags",
]
_F64_ARRAYS = ["node_forcing", "surf_inputs", "catch_inputs"]
_GROUP_FIELD = {"arc": "arc", "land": "land", "surf": "surf", "tank": "tank", "river": "river",
                "sewer": "sewer", "catchm": "catchm", "works": "works", "demand": "demand", "distrib": "distrib"}

_LIB = None
_LIB_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "csrc", "libwsi_b200.so")
if os.environ.get("WSI_B200_LIB"):  # development: a variant build of the same library (tools/build_variants.py)
    _LIB_PATH = os.path.abspath(os.environ["WSI_B200_LIB"])


class ExtensionMissingError(RuntimeError):
    pass


def library_path():
    return _LIB_PATH


def load_library():
    """Load ``libwsi_b200.so``; fail loudly when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(_LIB_PATH):
        raise ExtensionMissingError(
            f"{_LIB_PATH} not found: build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()'). wsi_b200 has no CPU fallback."
        )
    lib = C.CDLL(_LIB_PATH)
    lib.wsi_version.restype = C.c_int
    lib.wsi_last_error.restype = C.c_char_p
    lib.wsi_validate.argtypes = [C.POINTER(WsiModelDesc)]
    lib.wsi_validate.restype = C.c_int
    lib.wsi_num_blocks.argtypes = [C.POINTER(WsiModelDesc)]
    lib.wsi_num_blocks.restype = C.c_int
    lib.wsi_step.argtypes = [C.POINTER(WsiModelDesc), C.c_int64, C.c_void_p]
    lib.wsi_step.restype = C.c_int
    lib.wsi_run_steps.argtypes = [C.POINTER(WsiModelDesc), C.c_int64, C.c_int64, C.c_void_p]
    lib.wsi_run_steps.restype = C.c_int
    lib.wsi_sched_range.argtypes = [C.POINTER(WsiModelDesc), C.c_int64, C.c_int32, C.c_int32, C.c_uint32, C.c_void_p]
    lib.wsi_sched_range.restype = C.c_int
    lib.wsi_launch_count.restype = C.c_int64
    if lib.wsi_version() != ABI_VERSION:
        raise ExtensionMissingError("libwsi_b200.so ABI version mismatch; rebuild")
    lib.wsi_desc_size.restype = C.c_int
    if lib.wsi_desc_size() != C.sizeof(WsiModelDesc):
        raise ExtensionMissingError("libwsi_b200.so was built against another WsiModelDesc layout; rebuild")
    _LIB = lib
    return lib


def check(lib, rc):
    if rc != 0:
        msg = lib.wsi_last_error()
        raise RuntimeError(f"wsi_b200 error {rc}: {msg.decode() if msg else ''}")


MAX_CHAIN = 4  # kMaxChain in csrc/wsi_device.cuh: arcs a push may travel (kJunctionDepth junction Nodes)


def simple_chains(fm):
    """CSR ``(ptr, arcs)``: for every arc the arcs a push down it travels while each receiving node is
    a junction ``Node`` that forwards to exactly one arc (``nodes.py:62-66``: ``push_distributed`` with
    ``of_type = [Node, River, Waste, Reservoir]``; one out-arc and that arc eligible).  Empty when a
    junction on the way splits or dead-ends or the chain is longer than ``MAX_CHAIN`` -- such pushes run
    the general recursive code."""
    n_arcs = len(fm.arc_names)
    nrwr = L.FI["NRWR"]
    ptr = np.zeros(n_arcs + 1, dtype=np.int32)
    flat = []
    for a in range(n_arcs):
        arcs = [a]
        node = int(fm.arc_out[a])
        ok = True
        while int(fm.node_type[node]) == L.NT["Node"]:
            n_out = int(fm.out_ptr[node + 1] - fm.out_ptr[node])
            cb, ce = int(fm.cand_ptr[node * L.NF + nrwr]), int(fm.cand_ptr[node * L.NF + nrwr + 1])
            if n_out != 1 or ce - cb != 1 or len(arcs) >= MAX_CHAIN:
                ok = False
                break
            nxt = int(fm.out_arcs[fm.out_ptr[node]])
            arcs.append(nxt)
            node = int(fm.arc_out[nxt])
        if int(fm.node_type[node]) == L.NT["RiverReservoir"]:
            ok = False  # its push handlers look and push further downstream themselves (storage.py:1062-1107)
        if ok:
            flat.extend(arcs)
        ptr[a + 1] = len(flat)
    return ptr, np.array(flat or [0], dtype=np.int32)


def sched_split(fm):
    """Index of the first Land.run entry when something scheduled before it reaches into a Land's
    surfaces -- a Demand node with an arc to a Land (garden irrigation reads last step's evaporation
    and waters the garden before the surfaces run, demand.py:250-272, land.py:2402-2437) -- else 0."""
    land, demand = L.NT["Land"], L.NT["Demand"]
    if not any(int(fm.node_type[fm.arc_in[a]]) == demand and int(fm.node_type[fm.arc_out[a]]) == land
               for a in range(len(fm.arc_names))):
        return 0
    ops = [int(x) for x in fm.sched_op]
    return ops.index(L.OP["LAND_RUN"]) if L.OP["LAND_RUN"] in ops else 0


class ModelBuffers:
    """Owns the buffers of one flattened model on a device and its C descriptor."""

    STEP_BLOCK = 256  # replicas per block of wsi_step_kernel (kBlock in wsi_api.cu)
    OBJ_BLOCK = 128   # replicas per block of the object-parallel kernels (kObjBlock)

    def partial_rows(self):
        fm = self.fm
        items = len(fm.node_names) + len(fm.arc_names)
        if self.engine == 1:  # flat (item, replica) index
            return -(-(items * fm.R) // self.OBJ_BLOCK)
        return items * -(-fm.R // self.OBJ_BLOCK)

    def __init__(self, fm, device="cuda", probes=(), rec_t0=0, rec_T=None, record_mb=False, engine="auto",
                 reduce_probes=(), rec_members=None):
        from . import network

        self.net = network.plan_network(fm) if engine in ("auto", "network") else None
        if engine == "network" and self.net is None:
            raise ValueError("model is not eligible for the network engine: " + network.why_not(fm))
        if engine == "auto" and self.net is not None and fm.R * 8 > len(fm.node_names):
            self.net = None  # plenty of replicas: the replica engine has the parallelism
        self.engine = 1 if self.net is not None else 0
        # surface slots sorted by surface class (stable), for the per-class surface kernels
        types = fm.view("surf", "type")[0, :, 0].astype(np.int64) if fm.n_obj["surf"] else np.zeros(0, np.int64)
        order = np.argsort(types, kind="stable").astype(np.int32)
        # end_timestep items: every arc, decaying surfaces, tanks, rivers, sewers
        nd = fm.view("surf", "n_decay")[0].max(axis=1) if fm.n_obj["surf"] else np.zeros(0)
        items = [(0, a) for a in range(fm.n_obj["arc"])] + [(2, u) for u in range(fm.n_obj["surf"]) if nd[u] > 0]
        items += [(3, i) for i in range(fm.n_obj["tank"])] + [(4, i) for i in range(fm.n_obj["river"])]
        items += [(5, i) for i in range(fm.n_obj["sewer"])]
        items += [(6, i) for i in range(fm.n_obj.get("works", 0))] + [(7, i) for i in range(fm.n_obj.get("demand", 0))]
        items += [(8, i) for i in range(fm.n_obj.get("distrib", 0))]
        items += [(9, int(fm.node_gidx[i])) for i, t in enumerate(fm.node_type) if int(t) == L.NT["RiverReservoir"]]
        self.n_end = len(items)
        it = np.array(items or [(0, 0)], dtype=np.int32).reshape(-1, 2)
        fused = [int(t) in (L.NT["Land"], L.NT["Catchment"]) for t in fm.node_type]
        slot = np.full(len(fm.node_type), -1, dtype=np.int32)
        slot[np.nonzero(fused)[0]] = np.arange(int(np.sum(fused)), dtype=np.int32)
        self.n_fused = int(np.sum(fused))
        net = self.net or network.empty_plan(len(fm.node_names))
        net = dict(net, net_scratch=np.zeros(8, np.int32), net_done=np.zeros(max(len(net["net_item"]) // 8, 1) * fm.R, np.int32))
        self._extra_int = {k: np.ascontiguousarray(net[k], dtype=np.int32) for k in _NET_ARRAYS}
        chain_ptr, chain = simple_chains(fm)
        self._extra_int.update({"arc_chain_ptr": chain_ptr, "arc_chain": chain})
        self._extra_int.update({"surf_order": order if len(order) else np.zeros(1, np.int32), "mb_term_slot": slot,
                           "end_kind": np.ascontiguousarray(it[:, 0]), "end_obj": np.ascontiguousarray(it[:, 1])})
        self.surf_class_ptr = [int(np.searchsorted(types[order], k, side="left")) for k in range(4)] + [len(order)]
        self.fm = fm
        self.device = device
        self.probes = list(probes)
        # planes reduced over time on the device (running sum / maximum per member) and the window of
        # members whose full series are kept (None: all)
        self.reduce_probes = list(reduce_probes)
        self.rec_members = (0, 0) if rec_members is None else (int(rec_members[0]), int(rec_members[1]))
        if self.rec_members[1] and not (0 <= self.rec_members[0] and sum(self.rec_members) <= fm.R):
            raise ValueError("rec_members outside [0, R)")
        self.rec_R = self.rec_members[1] or fm.R
        self.rec_t0 = int(rec_t0)
        self.rec_T = int(fm.T if rec_T is None else rec_T)
        self.record_mb = bool(record_mb)
        self._keep = {}
        if device == "numpy":
            self._init_numpy()
        else:
            self._init_torch(device)
        self.desc = self._build_desc()

    def _all_groups(self):
        """Every group of the layout; builders that predate a group simply have no object of it."""
        fm = self.fm
        out = {}
        for g in L.GROUPS:
            if g in fm.groups:
                out[g] = fm.groups[g]
            else:
                fm.n_obj.setdefault(g, 0)
                out[g] = np.zeros((fm.n_fields(g), 0), dtype=np.float64)
        return out

    # ---------------------------------------------------------------- backends
    def _init_numpy(self):
        fm = self.fm
        k = self._keep
        for name in _INT_ARRAYS:
            src = self._extra_int[name] if name in self._extra_int else getattr(fm, name)
            k[name] = np.ascontiguousarray(src, dtype=np.int32)
        for name in _F64_ARRAYS:
            k[name] = np.ascontiguousarray(getattr(fm, name), dtype=np.float64)
        for g, arr in self._all_groups().items():
            k["group_" + g] = np.ascontiguousarray(arr, dtype=np.float64).copy()
        pr = np.array(self.probes or [(0, 0, 0)], dtype=np.int32).reshape(-1, 3)
        k["probe_group"] = np.ascontiguousarray(pr[:, 0])
        k["probe_plane"] = np.ascontiguousarray(pr[:, 1])
        k["probe_obj"] = np.ascontiguousarray(pr[:, 2])
        npb = max(len(self.probes), 1)
        k["out_probe"] = np.zeros((self.rec_T, npb, self.rec_R), dtype=np.float64)
        rd = np.array(self.reduce_probes or [(0, 0, 0)], dtype=np.int32).reshape(-1, 3)
        k["red_group"], k["red_plane"], k["red_obj"] = (np.ascontiguousarray(rd[:, i]) for i in range(3))
        k["red_out"] = np.zeros((2, max(len(self.reduce_probes), 1), fm.R), dtype=np.float64)
        k["red_out"][1] = -np.inf
        k["out_mb"] = (
            np.zeros((self.rec_T, len(fm.node_names), 3, fm.V, fm.R), dtype=np.float64)
            if self.record_mb
            else np.zeros(1, dtype=np.float64)
        )
        k["mb_partial"] = np.zeros((8, 3 * fm.V), dtype=np.float64)
        k["mb_terms"] = np.zeros((max(self.n_fused, 1), 3, fm.V, fm.R), dtype=np.float64)
        k["mb_sums"] = np.zeros((fm.T, 3, fm.V), dtype=np.float64)
        k["mb_count"] = np.zeros(fm.T, dtype=np.int32)
        k["node_flags"] = np.zeros(len(fm.node_names) * fm.R, dtype=np.uint32)
        self._ptr = lambda a: a.ctypes.data

    def _init_torch(self, device):
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("wsi_b200 needs a CUDA device (no CPU fallback on this path)")
        fm = self.fm
        k = self._keep
        dev = torch.device(device)
        self.torch_device = dev

        def up(a, dtype):
            return torch.from_numpy(np.ascontiguousarray(a)).to(dtype).to(dev).contiguous()

        for name in _INT_ARRAYS:
            k[name] = up(self._extra_int[name] if name in self._extra_int else getattr(fm, name), torch.int32)
        for name in _F64_ARRAYS:
            k[name] = up(getattr(fm, name), torch.float64)
        for g, arr in self._all_groups().items():
            k["group_" + g] = up(arr, torch.float64)
        pr = np.array(self.probes or [(0, 0, 0)], dtype=np.int32).reshape(-1, 3)
        k["probe_group"] = up(pr[:, 0], torch.int32)
        k["probe_plane"] = up(pr[:, 1], torch.int32)
        k["probe_obj"] = up(pr[:, 2], torch.int32)
        npb = max(len(self.probes), 1)
        k["out_probe"] = torch.zeros((self.rec_T, npb, self.rec_R), dtype=torch.float64, device=dev)
        rd = np.array(self.reduce_probes or [(0, 0, 0)], dtype=np.int32).reshape(-1, 3)
        k["red_group"], k["red_plane"], k["red_obj"] = (up(rd[:, i], torch.int32) for i in range(3))
        k["red_out"] = torch.zeros((2, max(len(self.reduce_probes), 1), fm.R), dtype=torch.float64, device=dev)
        k["red_out"][1] = float("-inf")
        k["out_mb"] = (
            torch.zeros((self.rec_T, len(fm.node_names), 3, fm.V, fm.R), dtype=torch.float64, device=dev)
            if self.record_mb
            else torch.zeros(1, dtype=torch.float64, device=dev)
        )
        k["mb_partial"] = torch.zeros((self.partial_rows(), 3 * fm.V), dtype=torch.float64, device=dev)
        k["mb_terms"] = torch.zeros((max(self.n_fused, 1), 3, fm.V, fm.R), dtype=torch.float64, device=dev)
        k["mb_sums"] = torch.zeros((fm.T, 3, fm.V), dtype=torch.float64, device=dev)
        k["mb_count"] = torch.zeros(fm.T, dtype=torch.int32, device=dev)
        # uint32 flags carried as int32 storage
        k["node_flags"] = torch.zeros(len(fm.node_names) * fm.R, dtype=torch.int32, device=dev)
        self._ptr = lambda a: a.data_ptr()

    # -------------------------------------------------------------- descriptor
    def _build_desc(self):
        fm = self.fm
        k = self._keep
        d = WsiModelDesc()
        d.abi_version = ABI_VERSION
        d.P, d.V, d.R, d.T = fm.P, fm.V, fm.R, fm.T
        d.n_nodes = len(fm.node_names)
        d.n_arcs = len(fm.arc_names)
        d.n_land = fm.n_obj["land"]
        d.n_surf = fm.n_obj["surf"]
        d.n_tank = fm.n_obj["tank"]
        d.n_river = fm.n_obj["river"]
        d.n_sewer = fm.n_obj["sewer"]
        d.n_catch = fm.n_obj["catchm"]
        d.n_works = fm.n_obj.get("works", 0)
        d.n_demand = fm.n_obj.get("demand", 0)
        d.n_distrib = fm.n_obj.get("distrib", 0)
        d.sched_split = sched_split(fm)
        rare = {L.NT[k] for k in ("WWTW", "FWTW", "Demand", "UnlimitedDistribution", "QueueGroundwater", "RiverReservoir", "Distribution")}
        pulls = any(int(t) in (L.NT["Reservoir"], L.NT["RiverReservoir"]) and fm.in_ptr[i + 1] > fm.in_ptr[i] for i, t in enumerate(fm.node_type))
        d.features = 1 if (pulls or any(int(t) in rare for t in fm.node_type)) else 0
        if fm.n_obj["surf"] and ((fm.view("surf", "irrigation_coefficient") >= 0).any() or (fm.view("surf", "climate_scale") != 1.0).any()):
            d.features |= 3  # bit 1: irrigation / variable-area surfaces (read by the surface kernel)
        d.n_sched = int(len(fm.sched_op))
        for i, v in enumerate(self.surf_class_ptr):
            d.surf_class_ptr[i] = v
        d.n_end = self.n_end
        net = self.net or {}
        d.engine = self.engine
        d.n_net_levels = int(len(net["net_level_ptr"]) - 1) if self.net else 0
        d.n_net_calc = int(net.get("n_net_calc", 0))
        d.n_net_final = int(net.get("n_net_final", 0))
        d.n_net_bnd = int(net.get("n_net_bnd", 0))
        d.nq = fm.nq
        d.n_fseries = fm.node_forcing.shape[1]
        d.n_months = fm.surf_inputs.shape[0]
        d.n_dep_series = fm.surf_inputs.shape[1]
        d.n_cseries = fm.catch_inputs.shape[1]
        d.n_probe = len(self.probes)
        d.rec_t0 = self.rec_t0
        d.rec_T = self.rec_T
        d.record_mb = 1 if self.record_mb else 0
        d.additive_mask = fm.additive_mask
        for i, v in enumerate(fm.idx):
            d.idx[i] = v
        d.float_accuracy = fm.float_accuracy
        P = self._ptr
        for name in _INT_ARRAYS:
            setattr(d, name, C.cast(P(k[name]), _i32p))
        for name in _F64_ARRAYS:
            setattr(d, name, C.cast(P(k[name]), _f64p))
        for g, fld in _GROUP_FIELD.items():
            setattr(d, fld, C.cast(P(k["group_" + g]), _f64p))
        d.n_red = len(self.reduce_probes)
        d.rec_r0, d.rec_R = self.rec_members
        d.red_out = C.cast(P(k["red_out"]), _f64p)
        for name in ("probe_group", "probe_plane", "probe_obj", "mb_count", "red_group", "red_plane", "red_obj"):
            setattr(d, name, C.cast(P(k[name]), _i32p))
        for name in ("out_probe", "out_mb", "mb_partial", "mb_sums", "mb_terms"):
            setattr(d, name, C.cast(P(k[name]), _f64p))
        d.node_flags = C.cast(P(k["node_flags"]), _u32p)
        self._level_ptr_host = np.ascontiguousarray(self._extra_int["net_level_ptr"], dtype=np.int32)
        d.net_level_ptr_host = self._level_ptr_host.ctypes.data_as(_i32p)
        return d

    # ----------------------------------------------------------------- access
    def group(self, name):
        """Group tensor/array ``[n_fields, n_obj*R]``."""
        return self._keep["group_" + name]

    def array(self, name):
        return self._keep[name]

    def to_numpy(self, name):
        a = self._keep[name]
        if self.device == "numpy":
            return a
        return a.detach().cpu().numpy()

    def groups_numpy(self):
        """Group planes as the model stands BETWEEN timesteps.  The replica engine does not zero the
        arcs at end_timestep (their transients expire by ``stamp``, csrc/wsi_device.cuh); what
        Arc.end_timestep leaves behind (arcs.py:259-265) is restored here."""
        out = {g: self.to_numpy("group_" + g) for g in self.fm.groups}
        if self.engine == 0 and out["arc"].size:
            a = np.array(out["arc"], copy=True)
            V = self.fm.V
            a[L.ARC.offset("flow_in", V)] = 0.0
            o = L.ARC.offset("vqip_in", V)
            a[o:o + V] = 0.0
            a[L.ARC.offset("stamp", V)] = -1.0
            out["arc"] = a
        return out

    def set_rec_window(self, rec_t0):
        """Move the recording window (out_probe row 0 == timestep rec_t0)."""
        self.rec_t0 = int(rec_t0)
        self.desc.rec_t0 = int(rec_t0)
