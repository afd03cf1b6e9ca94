"""Demand nodes (reference: wsimod/nodes/demand.py:14-324): parameters and per-step trackers.
``create_demand`` is schedule op DEMAND_CREATE on the device (csrc/wsi_works.cuh): the base class's
constant demand carrying ``pollutant_load`` (demand.py:119-131) and ResidentialDemand's house and
garden demands (demand.py:229-324)."""
from __future__ import annotations

from .nodes import Node


class Demand(Node):
    def __init__(self, name, constant_demand=0, pollutant_load=None, data_input_dict=None):
        self.constant_demand = constant_demand
        self.pollutant_load = {} if pollutant_load is None else pollutant_load
        super().__init__(name, data_input_dict={} if data_input_dict is None else data_input_dict)
        self.total_demand = self.empty_vqip()
        self.total_backup = self.empty_vqip()
        self.total_received = self.empty_vqip()

    def apply_overrides(self, overrides=None):
        overrides = overrides or {}
        self.constant_demand = overrides.pop("constant_demand", self.constant_demand)
        self.pollutant_load.update(overrides.pop("pollutant_load", {}))
        super().apply_overrides(overrides)


class ResidentialDemand(Demand):
    def __init__(self, population=1, pollutant_load=None, per_capita=0.12, gardening_efficiency=0.6 * 0.7,
                 data_input_dict=None, constant_temp=30, constant_weighting=0.2, **kwargs):
        self.gardening_efficiency = gardening_efficiency
        self.population = population
        self.per_capita = per_capita
        self.constant_weighting = constant_weighting
        self.constant_temp = constant_temp
        super().__init__(data_input_dict=data_input_dict, pollutant_load=pollutant_load, **kwargs)
        # the reference files it under "Demand" (demand.py:214)
        self.__class__.__name__ = "Demand"

    def apply_overrides(self, overrides=None):
        overrides = overrides or {}
        for key in ("gardening_efficiency", "population", "per_capita", "constant_weighting", "constant_temp"):
            setattr(self, key, overrides.pop(key, getattr(self, key)))
        super().apply_overrides(overrides)
