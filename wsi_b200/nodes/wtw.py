"""Treatment works (reference: wsimod/nodes/wtw.py:14-657): parameters and initial state of
``WTW`` / ``WWTW`` / ``FWTW``.  ``treat_water`` / ``calculate_discharge`` / ``make_discharge`` and the
push / pull handlers are schedule ops and protocol handlers on the device (csrc/wsi_works.cuh)."""
from __future__ import annotations

from ..core import constants
from .nodes import Node
from .tanks import Tank


class WTW(Node):
    def __init__(self, name, treatment_throughput_capacity=10, process_parameters=None, liquor_multiplier=None,
                 percent_solids=0.0002):
        self.treatment_throughput_capacity = treatment_throughput_capacity
        if process_parameters:
            self.process_parameters = process_parameters
        else:
            self.process_parameters = {x: {"constant": 0.01, "exponent": 1.001} for x in constants.ADDITIVE_POLLUTANTS}
        if liquor_multiplier:
            self._liquor_multiplier = liquor_multiplier
        else:
            self._liquor_multiplier = {x: 0.7 for x in constants.ADDITIVE_POLLUTANTS}
            self._liquor_multiplier["volume"] = 0.03
        self._percent_solids = percent_solids
        super().__init__(name)
        self.process_parameters["volume"] = {"constant": self.calculate_volume()}
        self.current_input = self.empty_vqip()
        self.treated = self.empty_vqip()
        self.liquor = self.empty_vqip()
        self.solids = self.empty_vqip()

    def calculate_volume(self):  # wtw.py:101-107
        return 1 - self._percent_solids - self._liquor_multiplier["volume"]

    @property
    def percent_solids(self):
        return self._percent_solids

    @percent_solids.setter
    def percent_solids(self, value):
        self._percent_solids = value
        self.process_parameters["volume"]["constant"] = self.calculate_volume()

    @property
    def liquor_multiplier(self):
        return self._liquor_multiplier

    @liquor_multiplier.setter
    def liquor_multiplier(self, value):
        self._liquor_multiplier.update(value)
        self.process_parameters["volume"]["constant"] = self.calculate_volume()

    def apply_overrides(self, overrides=None):
        overrides = overrides or {}
        self.percent_solids = overrides.pop("percent_solids", self._percent_solids)
        self.liquor_multiplier = overrides.pop("liquor_multiplier", self._liquor_multiplier)
        for key, value in overrides.pop("process_parameters", {}).items():
            self.process_parameters[key].update(value)
        self.treatment_throughput_capacity = overrides.pop("treatment_throughput_capacity",
                                                           self.treatment_throughput_capacity)
        super().apply_overrides(overrides)


class WWTW(WTW):
    def __init__(self, stormwater_storage_capacity=10, stormwater_storage_area=1, stormwater_storage_elevation=10,
                 **kwargs):
        self.stormwater_storage_capacity = stormwater_storage_capacity
        self.stormwater_storage_area = stormwater_storage_area
        self.stormwater_storage_elevation = stormwater_storage_elevation
        super().__init__(**kwargs)
        self.stormwater_tank = Tank(capacity=stormwater_storage_capacity, area=stormwater_storage_area,
                                    datum=stormwater_storage_elevation)
        self.liquor_ = self.empty_vqip()
        self.previous_input = self.empty_vqip()

    def apply_overrides(self, overrides=None):
        overrides = overrides or {}
        for key in ("stormwater_storage_capacity", "stormwater_storage_area", "stormwater_storage_elevation"):
            setattr(self, key, overrides.pop(key, getattr(self, key)))
        self.stormwater_tank.area = self.stormwater_storage_area
        self.stormwater_tank.capacity = self.stormwater_storage_capacity
        self.stormwater_tank.datum = self.stormwater_storage_elevation
        super().apply_overrides(overrides)

    def reinit(self):
        self.stormwater_tank.reinit()


class FWTW(WTW):
    def __init__(self, service_reservoir_storage_capacity=10, service_reservoir_storage_area=1,
                 service_reservoir_storage_elevation=10, service_reservoir_initial_storage=0, data_input_dict=None,
                 **kwargs):
        self.service_reservoir_storage_capacity = service_reservoir_storage_capacity
        self.service_reservoir_storage_area = service_reservoir_storage_area
        self.service_reservoir_storage_elevation = service_reservoir_storage_elevation
        self.service_reservoir_initial_storage = service_reservoir_initial_storage
        super().__init__(**kwargs)
        self.data_input_dict = {} if data_input_dict is None else data_input_dict
        self.total_deficit = self.empty_vqip()
        self.total_pulled = self.empty_vqip()
        self.previous_pulled = self.empty_vqip()
        self.unpushed_sludge = self.empty_vqip()
        self.service_reservoir_tank = Tank(capacity=service_reservoir_storage_capacity,
                                           area=service_reservoir_storage_area,
                                           datum=service_reservoir_storage_elevation,
                                           initial_storage=service_reservoir_initial_storage)

    def apply_overrides(self, overrides=None):
        overrides = overrides or {}
        for key in ("service_reservoir_storage_capacity", "service_reservoir_storage_area",
                    "service_reservoir_storage_elevation"):
            setattr(self, key, overrides.pop(key, getattr(self, key)))
        self.service_reservoir_tank.capacity = self.service_reservoir_storage_capacity
        self.service_reservoir_tank.area = self.service_reservoir_storage_area
        self.service_reservoir_tank.datum = self.service_reservoir_storage_elevation
        super().apply_overrides(overrides)

    def reinit(self):
        self.service_reservoir_tank.reinit()
