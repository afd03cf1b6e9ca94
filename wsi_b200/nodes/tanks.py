"""State/parameter holders with the reference's tank class names
(wsimod/nodes/tanks.py).  The store arithmetic itself runs on the device
(wsi_b200/csrc/wsi_device.cuh); these objects carry the initial state into the
flattener and receive the final state back after ``Model.run``."""
from __future__ import annotations

from ..core import constants
from ..core.core import DecayObj, WSIObj


class Tank(WSIObj):
    def __init__(self, capacity=0, area=1, datum=10, initial_storage=0):
        self.capacity = capacity
        self.area = area
        self.datum = datum
        self.initial_storage = initial_storage
        WSIObj.__init__(self)
        if isinstance(initial_storage, dict):
            self.storage = dict(initial_storage)
        else:  # a plain number is a volume without pollutants (tanks.py:43-50)
            self.storage = self.empty_vqip()
            self.storage["volume"] = initial_storage
        self.storage_ = dict(self.storage)

    def apply_overrides(self, overrides=None):
        overrides = overrides or {}
        for key in ("capacity", "area", "datum"):
            setattr(self, key, overrides.pop(key, getattr(self, key)))
        if overrides:
            print(f"No override behaviour defined for: {overrides.keys()}")

    def get_head(self, datum=None, non_head_storage=0):
        datum = self.datum if datum is None else datum
        return max(self.storage["volume"] - non_head_storage, 0) / self.area + datum

    def reinit(self):
        self.storage = self.empty_vqip()
        self.storage_ = self.empty_vqip()


class ResidenceTank(Tank):
    def __init__(self, residence_time=2, **kwargs):
        self.residence_time = residence_time
        super().__init__(**kwargs)

    def apply_overrides(self, overrides=None):
        overrides = overrides or {}
        self.residence_time = overrides.pop("residence_time", self.residence_time)
        super().apply_overrides(overrides)


class DecayTank(Tank, DecayObj):
    def __init__(self, decays=None, parent=None, **kwargs):
        self.parent = parent
        Tank.__init__(self, **kwargs)
        DecayObj.__init__(self, {} if decays is None else decays)

    def apply_overrides(self, overrides=None):
        overrides = overrides or {}
        self.decays.update(overrides.pop("decays", {}))
        super().apply_overrides(overrides)


class _InternalQueue:
    """The delay line of a QueueTank (reference: AltQueueArc, arcs.py:555-665): VQIPs keyed
    by remaining travel time."""

    def __init__(self, owner, number_of_timesteps=0):
        self.number_of_timesteps = number_of_timesteps
        self.queue = {0: owner.empty_vqip(), 1: owner.empty_vqip()}
        self.max_travel = 1


class QueueTank(Tank):
    def __init__(self, number_of_timesteps=0, **kwargs):
        self.number_of_timesteps = number_of_timesteps
        super().__init__(**kwargs)
        self.active_storage = dict(self.storage)
        self.internal_arc = _InternalQueue(self, number_of_timesteps)

    def apply_overrides(self, overrides=None):
        overrides = overrides or {}
        self.number_of_timesteps = overrides.pop("number_of_timesteps", self.number_of_timesteps)
        self.internal_arc.number_of_timesteps = self.number_of_timesteps
        super().apply_overrides(overrides)

    def reinit(self):
        super().reinit()
        self.active_storage = self.empty_vqip()
        self.internal_arc = _InternalQueue(self, self.number_of_timesteps)


class DecayQueueTank(QueueTank):
    """QueueTank whose delay line decays what enters it and what waits in it (reference:
    tanks.py:645-696 with DecayArcAlt, arcs.py:731-805); built-in delay of one timestep."""

    def __init__(self, decays=None, parent=None, number_of_timesteps=1, **kwargs):
        super().__init__(number_of_timesteps=number_of_timesteps, **kwargs)
        self.parent = parent
        self.internal_arc.decays = {} if decays is None else decays
        self.internal_arc.total_decayed = self.empty_vqip()

    def apply_overrides(self, overrides=None):
        overrides = overrides or {}
        self.internal_arc.decays.update(overrides.pop("decays", {}))
        super().apply_overrides(overrides)

    def reinit(self):
        decays = self.internal_arc.decays
        super().reinit()
        self.internal_arc.decays = decays
        self.internal_arc.total_decayed = self.empty_vqip()


__all__ = ["Tank", "ResidenceTank", "DecayTank", "QueueTank", "DecayQueueTank", "constants"]
