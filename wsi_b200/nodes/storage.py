"""Storage, Groundwater, River and Reservoir nodes (reference: wsimod/nodes/storage.py:15-245,
452-989).  Parameters and initial state only; ``distribute`` / ``infiltrate`` /
``calculate_discharge`` / ``make_abstractions`` are schedule ops on the device
(wsi_b200/csrc/wsi_nodes.cuh)."""
from __future__ import annotations

import warnings

from ..core import constants
from .nodes import Node
from .tanks import DecayQueueTank, DecayTank, QueueTank, Tank


class Storage(Node):
    def __init__(self, name, capacity=0, area=0, datum=0, decays=None, initial_storage=0, **kwargs):
        self.initial_storage = initial_storage
        self.capacity = capacity
        self.area = area
        self.datum = datum
        self.decays = decays
        super().__init__(name, **kwargs)
        tank_args = dict(capacity=capacity, area=area, datum=datum, initial_storage=initial_storage)
        self.tank = Tank(**tank_args) if decays is None else DecayTank(decays=decays, parent=self, **tank_args)

    def apply_overrides(self, overrides=None):
        overrides = overrides or {}
        for key in ("capacity", "area", "datum"):
            if key in overrides:
                setattr(self, key, overrides[key])
        if "decays" in overrides:
            if self.decays is None:
                raise ValueError("Attempting to override decays on a node initialised without decays")
            self.decays.update(overrides["decays"])
        self.tank.apply_overrides(overrides)
        super().apply_overrides(overrides)

    def get_percent(self):
        return self.tank.storage["volume"] / self.tank.capacity

    def reinit(self):
        self.tank.reinit()
        self.tank.storage["volume"] = self.initial_storage
        self.tank.storage_["volume"] = self.initial_storage


class Groundwater(Storage):
    def __init__(self, residence_time=200, infiltration_threshold=1, infiltration_pct=0, data_input_dict=None,
                 **kwargs):
        self.residence_time = residence_time
        self.infiltration_threshold = infiltration_threshold
        self.infiltration_pct = infiltration_pct
        super().__init__(**kwargs)
        # NOTE the reference's Node.__init__ resets data_input_dict to None after Groundwater
        # set it (storage.py:202-203, nodes.py:58); kept for parity
        self.data_input_dict = None

    def apply_overrides(self, overrides=None):
        overrides = overrides or {}
        for key in ("residence_time", "infiltration_threshold", "infiltration_pct"):
            setattr(self, key, overrides.pop(key, getattr(self, key)))
        super().apply_overrides(overrides)


class River(Storage):
    def __init__(self, depth=2, length=200, width=20, velocity=0.2 * constants.M_S_TO_M_DT, damp=0.1, mrf=0,
                 **kwargs):
        self.depth = depth
        if depth != 2:
            warnings.warn("the depth parameter is unused by River nodes (capacity is unbounded)")
        self.length = length
        self.width = width
        self.velocity = velocity
        self.damp = damp
        self.mrf = mrf
        kwargs["area"] = length * width
        kwargs["capacity"] = constants.UNBOUNDED_CAPACITY
        super().__init__(**kwargs)
        # in-river biochemistry parameters (storage.py:539-566)
        self.uptake_PNratio = 1 / 7.2
        self.bulk_density = 1300
        self.denpar_w = 0.0015
        self.T_wdays = 5
        self.halfsatINwater = 1.5 * constants.MG_L_TO_KG_M3
        self.hsatTP = 0.05 * constants.MG_L_TO_KG_M3
        self.limpppar = 0.1 * constants.MG_L_TO_KG_M3
        self.prodNpar = 0.001
        self.prodPpar = 0.0001
        self.muptNpar = 0.001
        self.muptPpar = 0.0001
        self.max_temp_lag = 20
        self.lagged_temperatures = []
        self.max_phosphorus_lag = 365
        self.lagged_total_phosphorus = []
        self.din_components = ["ammonia", "nitrate"]
        self.current_depth = 0
        self.bio_in = self.empty_vqip()
        self.bio_out = self.empty_vqip()

    _OVERRIDABLE = (
        "length", "width", "velocity", "damp", "mrf", "uptake_PNratio", "bulk_density", "denpar_w", "T_wdays",
        "halfsatINwater", "hsatTP", "limpppar", "prodNpar", "prodPpar", "muptNpar", "muptPpar", "max_temp_lag",
        "max_phosphorus_lag",
    )

    def apply_overrides(self, overrides=None):
        overrides = overrides or {}
        for key in self._OVERRIDABLE:
            if key in overrides:
                setattr(self, key, overrides.pop(key))
        overrides["area"] = self.length * self.width
        overrides["capacity"] = constants.UNBOUNDED_CAPACITY
        super().apply_overrides(overrides)


class Reservoir(Storage):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)


class RiverReservoir(Reservoir):
    """A reservoir on a river (reference: storage.py:992-1133): what it cannot hold spills on to the
    Node / River / Waste nodes downstream and counts towards ``environmental_flow``, which
    ``satisfy_environmental`` (an orchestration function) tops up.  Files itself under "Reservoir"."""

    def __init__(self, environmental_flow=0, **kwargs):
        self.environmental_flow = environmental_flow
        super().__init__(**kwargs)
        self.total_environmental_satisfied = 0
        self.__class__.__name__ = "Reservoir"

    def apply_overrides(self, overrides=None):
        overrides = overrides or {}
        self.environmental_flow = overrides.pop("environmental_flow", self.environmental_flow)
        super().apply_overrides(overrides)


class QueueGroundwater(Storage):
    """Groundwater whose baseflow follows a time-area diagram (reference: storage.py:248-449): pushes
    enter a QueueTank (a DecayQueueTank when ``decays`` are given) spread over ``timearea``;
    ``distribute`` releases what has arrived.  Like the reference it files itself under
    "Groundwater" and keeps answering push checks from the Tank its base class created first."""

    def __init__(self, timearea=None, data_input_dict=None, **kwargs):
        self.timearea = {0: 1} if timearea is None else timearea
        super().__init__(**kwargs)
        self.data_input_dict = {} if data_input_dict is None else data_input_dict
        self.__class__.__name__ = "Groundwater"
        self._stale_tank = self.tank
        args = dict(capacity=self.capacity, area=self.area, datum=self.datum, initial_storage=self.initial_storage)
        self.tank = QueueTank(**args) if self.decays is None else DecayQueueTank(decays=self.decays, parent=self, **args)

    def apply_overrides(self, overrides=None):
        overrides = overrides or {}
        self.timearea = overrides.pop("timearea", self.timearea)
        super().apply_overrides(overrides)
