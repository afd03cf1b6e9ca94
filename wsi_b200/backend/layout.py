"""Flat (structure-of-arrays) model layout shared by the CUDA library, the C oracle
and the golden fixtures.

A flattened WSIMOD model is a *template* graph (nodes, arcs, schedule) replicated
``R`` times (``R`` ensemble members / independent catchments, ``R == 1`` for an
ordinary model).  Every per-object quantity is one fp64 *plane* of length
``n_obj * R`` inside its group's 2-D tensor ``[n_fields, n_obj * R]``; object ``slot``
of replica ``r`` lives at column ``slot * R + r`` so that neighbouring threads
(replicas) touch neighbouring addresses (coalesced).

A VQIP (``wsimod/core/core.py:30``) is ``V = P + 1`` consecutive planes: plane 0 is
``volume``, planes ``1..P`` follow ``constants.POLLUTANTS`` order.

Field offsets are ``a + b * V``: all scalar fields first, VQIP fields last.  This
module is the single source of truth; ``tools/gen_layout.py`` turns it into
``include/wsi_layout.h`` (a test checks the committed header is current).
"""
from __future__ import annotations

from dataclasses import dataclass

MAX_DECAY = 4  # decaying pollutants per store (fixture uses 2: ammonia, nitrite)
MAX_STAGE = 16  # crop_factor_stages table length (fixture uses 9)
LAG_T = 21  # River.max_temp_lag + 1 (storage.py:560,753)
LAG_TP = 366  # River.max_phosphorus_lag + 1 (storage.py:563,758)
MAXP = 16  # compile-time bound on pollutants
COMPILED_P = (2, 12, 15)  # pollutant counts libwsi_b200.so has kernels for (csrc/wsi_cfg_*.cu); others are padded up
MAX_TIMEAREA = 8  # entries of a Sewer's pipe_timearea
ARC_DENY_PUSH, ARC_DENY_PULL = 1, 2  # arc_flags bits: PullArc / PushArc (arcs.py:808-893)
MAX_CAND = 8  # candidate out-arcs one push_distributed call may split over (kMaxCand in csrc/wsi_device.cuh)

# node types (Node.__class__.__name__ in the reference)
NODE_TYPES = [
    "Node",
    "Waste",
    "Land",
    "Sewer",
    "Groundwater",
    "Storage",
    "River",
    "Reservoir",
    "Catchment",
    "Boundary",  # not a reference class: the downstream end of a cut arc of a partitioned network --
    #              releases a prescribed VQIP series (totals) where the upstream river used to discharge
    "WWTW",  # wtw.py:165-420
    "FWTW",  # wtw.py:423-657
    "Demand",  # demand.py:14-324 (ResidentialDemand files itself under this name, demand.py:214)
    "UnlimitedDistribution",  # distribution.py:144-200: supplies whatever is pulled
    "QueueGroundwater",  # storage.py:248-449 (files itself under "Groundwater"): a QueueTank / DecayQueueTank store
    "RiverReservoir",  # storage.py:992-1133 (files itself under "Reservoir"): spills downstream, environmental flow
    "Distribution",  # distribution.py:101-141: conveys pulls, denies pushes, leaks a share to Groundwater
]
NT = {n: i for i, n in enumerate(NODE_TYPES)}

# surface types (land.py class names)
SURF_TYPES = ["Surface", "ImperviousSurface", "PerviousSurface", "GrowingSurface"]
ST = {n: i for i, n in enumerate(SURF_TYPES)}

# push filters == the ``of_type`` lists used by the reference call sites
FILTERS = {
    "ALL": None,  # of_type=None
    "SEWER": ["Sewer"],  # land.py:711, storage.py:240
    "GROUNDWATER": ["Groundwater"],  # land.py:211
    "RIVER_NODE": ["River", "Node"],  # land.py:225
    "NRWR": ["Node", "River", "Waste", "Reservoir"],  # nodes.py:63-71
    "NRW": ["Node", "River", "Waste"],  # storage.py:226,933 catchment.py:82
    "LAND": ["Land"],  # sewer.py:236
}
FILTER_NAMES = list(FILTERS.keys())
FI = {n: i for i, n in enumerate(FILTER_NAMES)}
NF = len(FILTER_NAMES)

# schedule ops (one orchestration entry applied to one node)
OPS = [
    "LAND_RUN",  # Land.run                    land.py:200
    "GW_INFILTRATE",  # Groundwater.infiltrate      storage.py:231
    "SEWER_DISCHARGE",  # Sewer.make_discharge        sewer.py:201
    "GW_DISTRIBUTE",  # Groundwater.distribute      storage.py:222
    "RIVER_CALC",  # River.calculate_discharge   storage.py:917
    "RESERVOIR_ABSTRACT",  # Reservoir.make_abstractions storage.py:983 (no in-arcs only)
    "CATCHMENT_ROUTE",  # Catchment.route             catchment.py:77
    "RIVER_DISTRIBUTE",  # River.distribute            storage.py:926
    "STORAGE_DISTRIBUTE",  # Storage.distribute          storage.py:120
    "NOP",  # e.g. Land.apply_irrigation without IrrigationSurface
    "BOUNDARY_RELEASE",  # partitioned networks: stands where the cut-off river's distribute stood
    "FWTW_TREAT",  # FWTW.treat_water            wtw.py:559
    "DEMAND_CREATE",  # Demand.create_demand        demand.py:79
    "WWTW_CALC",  # WWTW.calculate_discharge    wtw.py:278
    "WWTW_DISCHARGE",  # WWTW.make_discharge         wtw.py:306
    "QGW_DISTRIBUTE",  # QueueGroundwater.distribute storage.py:345-363
    "RESERVOIR_ENVIRONMENTAL",  # RiverReservoir.satisfy_environmental storage.py:1109-1127
    "LAND_IRRIGATE",  # Land.apply_irrigation land.py:192-198 -> IrrigationSurface.irrigate land.py:2347-2381
]
OP = {n: i for i, n in enumerate(OPS)}


@dataclass(frozen=True)
class Field:
    name: str
    n: int = 1  # number of planes (scalars/arrays); for vqip fields: number of VQIPs
    vqip: bool = False


class Group:
    """Ordered field list -> offsets ``a + b*V``."""

    def __init__(self, name: str, fields: list[Field]):
        self.name = name
        self.fields = fields
        self.scalars = [f for f in fields if not f.vqip]
        self.vqips = [f for f in fields if f.vqip]
        self.off: dict[str, tuple[int, int]] = {}
        a = 0
        for f in self.scalars:
            self.off[f.name] = (a, 0)
            a += f.n
        self.n_scalar = a
        b = 0
        for f in self.vqips:
            self.off[f.name] = (a, b)
            b += f.n
        self.n_vqip = b

    def offset(self, name: str, V: int) -> int:
        a, b = self.off[name]
        return a + b * V

    def n_fields(self, V: int) -> int:
        return self.n_scalar + self.n_vqip * V

    def size(self, name: str, V: int) -> int:
        for f in self.fields:
            if f.name == name:
                return f.n * (V if f.vqip else 1)
        raise KeyError(name)


def S(name, n=1):
    return Field(name, n, False)


def Q(name, n=1):
    return Field(name, n, True)


_DECAY = [
    S("n_decay"),
    S("decay_pol", MAX_DECAY),  # vqip index (1..P) of the decaying pollutant
    S("decay_const", MAX_DECAY),
    S("decay_exp", MAX_DECAY),
    S("total_decayed", MAX_DECAY),  # state: DecayObj.total_decayed (core.py:472)
    S("pre_decay", MAX_DECAY),  # state: storage_[pol] before decay (tanks.py:445)
]

ARC = Group(
    "arc",
    [
        S("capacity"),
        S("preference"),
        S("flow_in"),  # per-step transient (arcs.py:64,152)
        S("pending_mode"),  # network engine: 0 = the sender already booked `pending` on the arc and
        #                     the receiver only adds it to its tank; 1 = the receiver runs the whole push
        S("stamp"),  # replica engine: the timestep flow_in / vqip_in belong to (anything else reads as the
        #              empty arc of arcs.py:259-265, so end_timestep never has to zero the arcs); -1 = none
        Q("vqip_in"),  # per-step transient; vqip_out aliases it (arcs.py:155-156)
        Q("pending"),  # network engine: the release the arc's sender parked for its receiver
    ],
)

LAND = Group(
    "land",
    [
        S("res_surface"),  # surface_residence_time
        S("res_subsurface"),
        S("res_percolation"),
        S("forcing_id"),
        S("node"),  # node slot of this Land (status flags)
        S("garden_surf"),  # surface slot of the Land's GardenSurface (-1: none): it answers pushes tagged
        #                    ("Demand", "Garden") (land.py:120-128, 2384-2437)
        Q("surface_runoff"),  # ResidenceTank storages (land.py:138-149)
        Q("subsurface_runoff"),
        Q("percolation"),
    ],
)

SURF = Group(
    "surf",
    [
        S("type"),
        S("land"),  # index of the parent Land in the land group
        S("area"),
        S("capacity"),
        S("depth"),
        S("has_load"),
        S("has_dep"),
        S("dep_id"),
        *_DECAY,
        # ImperviousSurface (land.py:587)
        S("et0_to_e"),
        # PerviousSurface (land.py:727)
        S("field_capacity_m"),
        S("wilting_point_m"),
        S("infiltration_capacity"),
        S("surface_coefficient"),
        S("percolation_coefficient"),
        S("subsurface_coefficient"),
        S("ihacres_p"),
        S("et0_coefficient"),  # parameter for Pervious, state for Growing
        S("soil_temp_w_prev"),
        S("soil_temp_w_air"),
        S("soil_temp_w_deep"),
        S("soil_temp_deep"),
        # GrowingSurface parameters (land.py:1085-1270)
        S("rooting_depth"),
        S("ET_depletion_factor"),
        S("total_available_water"),
        S("readily_available_water"),
        S("sowing_day"),
        S("harvest_day"),
        S("autumn_sow"),
        S("n_stage"),
        S("stage_dates", MAX_STAGE),
        S("stage_vals", MAX_STAGE),
        S("hs_calendar", 5),  # harvest_sow_calendar
        S("ground_cover_stages", 5),
        S("crop_cover_stages", 5),
        S("satact"),
        S("thetaupp"),
        S("thetalow"),
        S("thetapow"),
        S("uptake1"),
        S("uptake2"),
        S("uptake3"),
        S("uptake_PNratio"),
        S("erodibility"),
        S("sreroexp"),
        S("cohesion"),
        S("slope"),
        S("srfilt"),
        S("macrofilt"),
        S("limpar"),
        S("exppar"),
        S("hsatINs"),
        S("denpar"),
        S("adsorption_nr_limit"),
        S("adsorption_nr_maxiter"),
        S("kfr"),
        S("nfr"),
        S("kadsdes"),
        S("bulk_density"),
        # NutrientPool parameters (nutrient_pool.py:14-24), [N, P]
        S("frac_dry_n_to_dis_inorg"),
        S("frac_dry_n_to_fast"),
        S("frac_manure_to_dis_inorg", 2),
        S("frac_manure_to_fast", 2),
        S("degrhpar", 2),
        S("dishpar", 2),
        S("minfpar", 2),
        S("disfpar", 2),
        S("immobdpar", 2),
        # GrowingSurface state
        S("days_after_sow"),  # -1 == None
        S("crop_cover"),
        S("ground_cover"),
        S("crop_factor"),
        S("pool_fast", 2),  # NutrientStore storages [N, P]
        S("pool_humus", 2),
        S("pool_dis_inorg", 2),
        S("pool_dis_org", 2),
        S("pool_ads_inorg", 2),
        # per-step diagnostics (Model.run surfaces record, model.py:1081-1118)
        S("out_percolation"),
        S("out_subsurface"),
        S("out_surface"),
        S("out_evaporation"),
        S("out_precipitation"),
        S("out_tank_recharge"),
        S("irrigated_at"),  # GardenSurface: timestep whose irrigation arrived before the surface ran (-1: none)
        S("irrigation_coefficient"),  # IrrigationSurface (land.py:2317-2329); -1: not one
        S("climate_scale"),  # VariableAreaSurface: current_surface_area / area (land.py:2449-2461); 1 otherwise
        Q("storage"),
        Q("load"),  # pollutant_load (plane 0 unused)
        # per-step exchange between the surface kernels and the land phase (scratch)
        Q("flow_a"),  # infiltration_excess (pervious) / ponded runoff (impervious)
        Q("flow_b"),  # subsurface_flow
        Q("flow_c"),  # percolation
        Q("mb_in"),  # this surface's share of running_inflow_mb (volume + additive lanes)
        Q("mb_out"),  # ... of running_outflow_mb
        Q("mb_ds"),  # surface.ds() (tanks.py:449-457)
        Q("pre_irrigation"),  # GardenSurface: storage before this step's irrigation (= storage_ for ds())
    ],
)

TANK = Group(
    "tank",
    [
        S("capacity"),
        S("area"),
        S("residence_time"),
        S("infiltration_threshold"),
        S("infiltration_pct"),
        S("forcing_id"),
        S("is_decay"),  # DecayTank vs Tank (storage.py:53-68)
        S("environmental_flow"),  # RiverReservoir (storage.py:995, 1031)
        S("env_satisfied"),  # state: total_environmental_satisfied (storage.py:1035), reset at end_timestep
        *_DECAY,
        Q("storage"),
        Q("storage_"),  # lagged storage for mass balance (tanks.py:40,351)
    ],
)

RIVER = Group(
    "river",
    [
        S("tank"),  # index into the tank group
        S("riverrc"),  # precomputed on the host (storage.py:900-915)
        S("area"),
        S("uptake_PNratio"),
        S("denpar_w"),
        S("T_wdays"),
        S("halfsatINwater"),
        S("hsatTP"),
        S("limpppar"),
        S("prodNpar"),
        S("prodPpar"),
        S("muptNpar"),
        S("muptPpar"),
        S("max_temp_lag"),
        S("max_phosphorus_lag"),
        S("mrf"),  # minimum required flow, volume per timestep (storage.py:477, 676-680)
        S("n_lag_T"),  # state: len(lagged_temperatures)
        S("head_T"),  # state: ring index of the oldest entry
        S("n_lag_TP"),
        S("head_TP"),
        S("lag_T", LAG_T),
        S("lag_TP", LAG_TP),
        Q("bio_in"),  # per-step transient (storage.py:592-596)
        Q("bio_out"),
    ],
)

SEWER = Group(
    "sewer",
    [
        S("capacity"),
        S("pipe_time"),
        S("qmask"),  # state: which travel-time keys exist in internal_arc.queue
        S("ia_flow_in"),  # per-step transient: internal_arc.flow_in (arcs.py:427)
        S("n_timearea"),  # pipe_timearea (sewer.py:62-66): pushes tagged "Land"/"Demand" spread over travel times
        S("ta_time", MAX_TIMEAREA),
        S("ta_frac", MAX_TIMEAREA),
        S("kind"),  # 0 Sewer (sewer.py), 1 QueueGroundwater (storage.py:248)
        S("queue_delay"),  # internal_arc.number_of_timesteps, added to every travel time (arcs.py:385)
        S("stale_capacity"),  # QueueGroundwater answers push checks from the Tank its Storage base class
        S("stale_volume"),  # created first and no longer updates (storage.py:60, 301-315)
        S("forcing_id"),  # temperature series for the queue's decays
        S("n_decay"),  # DecayQueueTank / DecayArcAlt (tanks.py:645-696, arcs.py:731-805)
        S("decay_pol", MAX_DECAY),
        S("decay_const", MAX_DECAY),
        S("decay_exp", MAX_DECAY),
        S("total_decayed", MAX_DECAY),  # state: internal_arc.total_decayed
        Q("storage"),
        Q("storage_"),
        Q("active_storage"),
        # followed by NQ queue VQIPs; NQ is a model constant, so the queue is
        # addressed as offset("queue") + k*V
        Q("queue"),
    ],
)

CATCH = Group("catchm", [S("cid")])

# WWTW / FWTW (wtw.py): the WTW process model plus the works' own tank (stormwater tank / service
# reservoir: an object of the tank group)
WORKS = Group(
    "works",
    [
        S("kind"),  # 0 WWTW, 1 FWTW
        S("tank"),  # index into the tank group
        S("throughput_capacity"),  # treatment_throughput_capacity
        S("percent_solids"),
        Q("liquor_mult"),  # liquor_multiplier: plane 0 = volume, additive planes per pollutant
        Q("proc_const"),  # process_parameters[...]["constant"]: plane 0 = volume (wtw.py:101-107)
        Q("proc_exp"),  # process_parameters[...]["exponent"] (plane 0 unused)
        Q("current_input"),  # state (wtw.py:92-95)
        Q("treated"),
        Q("liquor"),
        Q("solids"),
        Q("liquor_"),  # WWTW: liquor at the end of the previous step (mass balance)
        Q("total_deficit"),  # FWTW (wtw.py:512-515)
        Q("total_pulled"),
        Q("previous_pulled"),
        Q("unpushed_sludge"),
    ],
)

# ResidentialDemand (demand.py:153-324)
DEMAND = Group(
    "demand",
    [
        S("population"),
        S("per_capita"),
        S("gardening_efficiency"),
        S("constant_weighting"),
        S("constant_temp"),
        S("forcing_id"),  # temperature series
        Q("pollutant_load"),  # per person and timestep (plane 0 unused)
        Q("total_demand"),  # per-step trackers (mass balance, demand.py:57-64)
        Q("total_backup"),
        Q("total_received"),
    ],
)

# UnlimitedDistribution (distribution.py:144-200)
DISTRIB = Group("distrib", [S("leakage"), Q("supplied")])  # leakage: plain Distribution (distribution.py:101-141)

GROUPS = {g.name: g for g in (ARC, LAND, SURF, TANK, RIVER, SEWER, CATCH, WORKS, DEMAND, DISTRIB)}
GROUP_ID = {n: i for i, n in enumerate(GROUPS)}

# surface monthly data columns (land.py:540-569,1659-1691)
DEP_COLS = [
    "nhx-dry",
    "noy-dry",
    "srp-dry",
    "nhx-wet",
    "noy-wet",
    "srp-wet",
    "nhx-fertiliser",
    "noy-fertiliser",
    "srp-fertiliser",
    "nhx-manure",
    "noy-manure",
    "srp-manure",
]
N_DEP = len(DEP_COLS)

# named pollutants the kernels special-case (membership tests land.py:427,1276)
NAMED = [
    "nitrate",
    "nitrite",
    "ammonia",
    "phosphate",
    "org-phosphorus",
    "org-nitrogen",
    "solids",
    "temperature",
]

# per-node mass-balance record: (in, ds, out) VQIPs
MB_TERMS = 3


def sewer_n_fields(V: int, nq: int) -> int:
    return SEWER.n_scalar + (SEWER.n_vqip - 1 + nq) * V


def c_header() -> str:
    """Render include/wsi_layout.h."""
    out = []
    w = out.append
    w("/* GENERATED by tools/gen_layout.py from wsi_b200/backend/layout.py -- do not edit. */")
    w("#ifndef WSI_LAYOUT_H")
    w("#define WSI_LAYOUT_H")
    w("")
    for k, v in (
        ("WSI_MAX_DECAY", MAX_DECAY),
        ("WSI_MAX_STAGE", MAX_STAGE),
        ("WSI_LAG_T", LAG_T),
        ("WSI_LAG_TP", LAG_TP),
        ("WSI_MAXP", MAXP),
        ("WSI_MAX_CAND", MAX_CAND),
        ("WSI_MAX_TIMEAREA", MAX_TIMEAREA),
        ("WSI_NF", NF),
        ("WSI_N_DEP", N_DEP),
    ):
        w(f"#define {k} {v}")
    w("")
    w("/* node types */")
    for n, i in NT.items():
        w(f"#define WSI_NT_{n.upper()} {i}")
    w("/* surface types */")
    for n, i in ST.items():
        w(f"#define WSI_ST_{n.upper()} {i}")
    w("/* push filters (of_type lists) */")
    for n, i in FI.items():
        w(f"#define WSI_FILTER_{n} {i}")
    w("/* schedule ops */")
    for n, i in OP.items():
        w(f"#define WSI_OP_{n} {i}")
    w("/* surface monthly data columns */")
    for i, n in enumerate(DEP_COLS):
        w(f"#define WSI_DEP_{n.upper().replace('-', '_')} {i}")
    w("")
    for g in GROUPS.values():
        G = g.name.upper()
        w(f"/* group {g.name}: field offsets are a + b*V planes */")
        for f in g.fields:
            a, b = g.off[f.name]
            nm = f.name.upper().rstrip("_") + ("_LAG" if f.name.endswith("_") else "")
            if f.vqip:
                w(f"#define WSI_{G}_{nm}(V) ({a} + {b} * (V))")
            else:
                w(f"#define WSI_{G}_{nm} {a}")
        w(f"#define WSI_{G}_NSCALAR {g.n_scalar}")
        w(f"#define WSI_{G}_NVQIP {g.n_vqip}")
        w("")
    w("#endif")
    return "\n".join(out) + "\n"
