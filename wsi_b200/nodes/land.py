"""Land node and its surfaces (reference: wsimod/nodes/land.py).

These classes reproduce the reference's constructor signatures, derived parameters and
initial state so that YAML model files and user scripts build the same model; the
hydrology / nutrient equations themselves are the fused surface kernel
(wsi_b200/csrc/wsi_land.cuh)."""
from __future__ import annotations

import sys

from ..core import constants
from .nodes import Node
from .nutrient_pool import NutrientPool
from .tanks import DecayTank, ResidenceTank


class Land(Node):
    def __init__(self, name, subsurface_residence_time=5, percolation_residence_time=50, surface_residence_time=1,
                 surfaces=None, data_input_dict=None):
        self.subsurface_residence_time = subsurface_residence_time
        self.percolation_residence_time = percolation_residence_time
        self.surface_residence_time = surface_residence_time
        data_input_dict = {} if data_input_dict is None else data_input_dict
        super().__init__(name, data_input_dict=data_input_dict)
        self.surfaces = []
        for spec in list(surfaces or []):
            spec = dict(spec)
            cls = getattr(sys.modules[__name__], spec.pop("type_"))
            self.surfaces.append(cls(parent=self, **spec))
        unbounded = constants.UNBOUNDED_CAPACITY
        self.subsurface_runoff = ResidenceTank(residence_time=subsurface_residence_time, capacity=unbounded)
        self.percolation = ResidenceTank(residence_time=percolation_residence_time, capacity=unbounded)
        self.surface_runoff = ResidenceTank(residence_time=surface_residence_time, capacity=unbounded)
        self.running_inflow_mb = self.empty_vqip()
        self.running_outflow_mb = self.empty_vqip()

    def apply_overrides(self, overrides=None):
        overrides = overrides or {}
        for key, tank in (("surface_residence_time", self.surface_runoff),
                          ("subsurface_residence_time", self.subsurface_runoff),
                          ("percolation_residence_time", self.percolation)):
            setattr(self, key, overrides.pop(key, getattr(self, key)))
            tank.residence_time = getattr(self, key)
        for surface, override in overrides.pop("surfaces", {}).items():
            self.get_surface(surface).apply_overrides(override)
        super().apply_overrides(overrides)

    def get_surface(self, surface_):
        for surface in self.surfaces:
            if surface.surface == surface_:
                return surface
        return None


class Surface(DecayTank):
    def __init__(self, surface="", area=1, depth=1, data_input_dict=None, pollutant_load=None, **kwargs):
        self.depth = depth
        self.data_input_dict = {} if data_input_dict is None else data_input_dict
        self.surface = surface
        self.pollutant_load = {} if pollutant_load is None else pollutant_load
        super().__init__(capacity=area * depth, area=area, **kwargs)

    def apply_overrides(self, overrides=None):
        overrides = overrides or {}
        self.surface = overrides.pop("surface", self.surface)
        self.pollutant_load.update(overrides.pop("pollutant_load", {}))
        self.area = overrides.pop("area", self.area)
        self.depth = overrides.pop("depth", self.depth)
        self.capacity = self.area * self.depth
        if "capacity" in overrides:
            overrides.pop("capacity")
            print("Warning: capacity overrides are ignored for surfaces; give depth and area")
        content = overrides.pop("data_input_dict", self.data_input_dict)
        if isinstance(content, str):
            from ..orchestration.model import read_csv

            self.data_input_dict = read_csv(content)
        elif isinstance(content, dict):
            self.data_input_dict = content
        else:
            raise ValueError(f"{content.__class__} is not a recognised format for data_input_dict")
        super().apply_overrides(overrides)


class ImperviousSurface(Surface):
    def __init__(self, pore_depth=0, et0_to_e=1, **kwargs):
        self.et0_to_e = et0_to_e
        self.pore_depth = pore_depth
        super().__init__(depth=pore_depth, **kwargs)
        self.evaporation = self.empty_vqip()
        self.precipitation = self.empty_vqip()

    def apply_overrides(self, overrides=None):
        overrides = overrides or {}
        self.et0_to_e = overrides.pop("et0_to_e", self.et0_to_e)
        if "depth" in overrides:
            overrides.pop("depth")
            print("ERROR: give pore_depth, not depth, for an ImperviousSurface")
        self.pore_depth = overrides.pop("pore_depth", self.pore_depth)
        overrides["depth"] = self.pore_depth
        super().apply_overrides(overrides)


class PerviousSurface(Surface):
    def __init__(self, depth=0.75, total_porosity=0.4, field_capacity=0.3, wilting_point=0.12,
                 infiltration_capacity=0.5, surface_coefficient=0.05, percolation_coefficient=0.75,
                 et0_coefficient=0.5, ihacres_p=10, **kwargs):
        self.field_capacity = field_capacity
        self.field_capacity_m = field_capacity * depth
        self.wilting_point = wilting_point
        self.wilting_point_m = wilting_point * depth
        self.infiltration_capacity = infiltration_capacity
        self.surface_coefficient = surface_coefficient
        self.percolation_coefficient = percolation_coefficient
        self.et0_coefficient = et0_coefficient
        self.ihacres_p = ihacres_p
        self.total_porosity = total_porosity
        self.soil_temp_w_prev = 0.1
        self.soil_temp_w_air = 0.6
        self.soil_temp_w_deep = 0.1
        self.soil_temp_deep = 10
        super().__init__(depth=depth * total_porosity, **kwargs)
        self.subsurface_coefficient = 1 - self.percolation_coefficient
        self.infiltration_excess = self.empty_vqip()
        self.subsurface_flow = self.empty_vqip()
        self.percolation = self.empty_vqip()
        self.tank_recharge = 0
        self.evaporation = self.empty_vqip()
        self.precipitation = self.empty_vqip()

    _OVERRIDABLE = (
        "field_capacity", "wilting_point", "total_porosity", "infiltration_capacity", "surface_coefficient",
        "percolation_coefficient", "et0_coefficient", "ihacres_p", "soil_temp_w_prev", "soil_temp_w_air",
        "soil_temp_w_deep", "soil_temp_deep",
    )

    def apply_overrides(self, overrides=None):
        overrides = overrides or {}
        self.depth /= self.total_porosity  # back to the physical (root) depth
        for key in self._OVERRIDABLE:
            setattr(self, key, overrides.pop(key, getattr(self, key)))
        self.subsurface_coefficient = 1 - self.percolation_coefficient
        super().apply_overrides(overrides)
        self.field_capacity_m = self.field_capacity * self.depth
        self.wilting_point_m = self.wilting_point * self.depth
        self.depth *= self.total_porosity
        self.capacity = self.depth * self.area


class GrowingSurface(PerviousSurface):
    def __init__(self, rooting_depth=1, ET_depletion_factor=1, crop_factor_stages=None,
                 crop_factor_stage_dates=None, sowing_day=1, harvest_day=365, initial_soil_storage=None, **kwargs):
        self.ET_depletion_factor = ET_depletion_factor
        self.rooting_depth = rooting_depth
        self.crop_cover_max = 0.9
        self.ground_cover_max = 0.3
        self.crop_factor_stages = [1, 1] if crop_factor_stages is None else crop_factor_stages
        self.crop_factor_stage_dates = [0, 365] if crop_factor_stage_dates is None else crop_factor_stage_dates
        self.sowing_day = sowing_day
        self.harvest_day = harvest_day
        # HYPE-derived constants (land.py:1204-1250)
        self.satact, self.thetaupp, self.thetalow, self.thetapow = 0.6, 0.12, 0.08, 1
        self.uptake1, self.uptake2, self.uptake3, self.uptake_PNratio = 15, 1, 0.02, 1 / 7.2
        self.erodibility, self.sreroexp, self.cohesion, self.slope = 0.0025, 1.2, 1, 5
        self.srfilt, self.macrofilt = 0.7, 0.01
        self.limpar, self.exppar, self.hsatINs, self.denpar = 0.7, 2.5, 1, 0.015
        self.adosorption_nr_limit = 0.00001  # sic, the reference's spelling
        self.adsorption_nr_maxiter = 20
        self.kfr, self.nfr, self.kadsdes = 153.7, 1 / 2.6, 0.03
        self.bulk_density = 1300
        super().__init__(depth=rooting_depth, **kwargs)
        self._refresh_calendar()
        self.days_after_sow = None
        self.crop_cover = 0
        self.ground_cover = 0
        self.crop_factor = 0
        self.et0_coefficient = 1
        self.nutrient_pool = NutrientPool()
        if "nitrate" in constants.POLLUTANTS:
            st, pool = self.storage, self.nutrient_pool
            pool.dissolved_inorganic_pool.storage["P"] = st["phosphate"]
            pool.dissolved_inorganic_pool.storage["N"] = st["nitrate"] + st["ammonia"] + st["nitrite"]
            pool.dissolved_organic_pool.storage["P"] = st["org-phosphorus"]
            pool.dissolved_organic_pool.storage["N"] = st["org-nitrogen"]
            if initial_soil_storage:
                self.initial_soil_storage = initial_soil_storage
                s = initial_soil_storage
                pool.adsorbed_inorganic_pool.storage["P"] = s["phosphate"]
                pool.adsorbed_inorganic_pool.storage["N"] = s["ammonia"] + s["nitrate"] + s["nitrite"]
                pool.fast_pool.storage["N"] = s["org-nitrogen"]
                pool.fast_pool.storage["P"] = s["org-phosphorus"]

    def _refresh_calendar(self):
        # infer_sow_harvest_calendar / calculate_available_water (land.py:1326-1365)
        self.harvest_sow_calendar = [0, self.sowing_day, self.harvest_day, self.harvest_day + 1, 365]
        self.ground_cover_stages = [0, 0, self.ground_cover_max, 0, 0]
        self.crop_cover_stages = [0, 0, self.crop_cover_max, 0, 0]
        self.autumn_sow = self.sowing_day > 181
        self.total_available_water = self.field_capacity_m - self.wilting_point_m
        if self.total_available_water < 0:
            print("warning: TAW < 0...")
        self.readily_available_water = self.total_available_water * self.ET_depletion_factor

    _GROW_OVERRIDABLE = (
        "ET_depletion_factor", "crop_cover_max", "ground_cover_max", "crop_factor_stages",
        "crop_factor_stage_dates", "sowing_day", "harvest_day", "satact", "thetaupp", "thetalow", "thetapow",
        "uptake1", "uptake2", "uptake3", "uptake_PNratio", "erodibility", "sreroexp", "cohesion", "slope", "srfilt",
        "macrofilt", "limpar", "exppar", "hsatINs", "denpar", "adosorption_nr_limit", "adsorption_nr_maxiter", "kfr",
        "nfr", "kadsdes", "bulk_density",
    )

    def apply_overrides(self, overrides=None):
        overrides = overrides or {}
        for key in self._GROW_OVERRIDABLE:
            setattr(self, key, overrides.pop(key, getattr(self, key)))
        if "depth" in overrides:
            overrides.pop("depth")
            print("ERROR: give rooting_depth, not depth, for a GrowingSurface")
        self.rooting_depth = overrides.pop("rooting_depth", self.rooting_depth)
        overrides["depth"] = self.rooting_depth
        super().apply_overrides(overrides)
        self._refresh_calendar()


class IrrigationSurface(GrowingSurface):
    """A GrowingSurface whose crop water demand not met by rain is pulled through the parent Land's
    in-arcs after the surfaces have run (reference: land.py:2314-2381; the kernels: land_irrigate in
    csrc/wsi_works.cuh, scheduled by the Land's ``apply_irrigation`` orchestration entry)."""

    def __init__(self, irrigation_coefficient=0.1, **kwargs):
        self.irrigation_coefficient = irrigation_coefficient
        super().__init__(**kwargs)


class GardenSurface(GrowingSurface):
    """Behaves as a GrowingSurface unless a Demand node pushes irrigation (land.py:2384-2437)."""


class VariableAreaSurface(GrowingSurface):
    def __init__(self, current_surface_area=0, **kwargs):
        super().__init__(**kwargs)
        self.current_surface_area = current_surface_area
