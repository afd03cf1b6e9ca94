"""Sewer node (reference: wsimod/nodes/sewer.py:14-248): a QueueTank with a pipe travel
time.  ``make_discharge`` is schedule op SEWER_DISCHARGE on the device."""
from __future__ import annotations

from .nodes import Node
from .tanks import QueueTank


class Sewer(Node):
    def __init__(self, name, capacity=0, pipe_time=0, pipe_timearea=None, chamber_area=1, chamber_floor=10,
                 data_input_dict=None):
        self.capacity = capacity
        self.pipe_time = pipe_time
        self.pipe_timearea = {0: 1} if pipe_timearea is None else pipe_timearea
        self.chamber_area = chamber_area
        self.chamber_floor = chamber_floor
        super().__init__(name)
        self.data_input_dict = {} if data_input_dict is None else data_input_dict
        self.sewer_tank = QueueTank(capacity=capacity, number_of_timesteps=0, datum=chamber_floor, area=chamber_area)

    def apply_overrides(self, overrides=None):
        overrides = overrides or {}
        for key in ("capacity", "chamber_area", "chamber_floor", "pipe_time", "pipe_timearea"):
            setattr(self, key, overrides.pop(key, getattr(self, key)))
        self.sewer_tank.capacity = self.capacity
        self.sewer_tank.area = self.chamber_area
        self.sewer_tank.datum = self.chamber_floor
        super().apply_overrides(overrides)

    def reinit(self):
        self.sewer_tank.reinit()
