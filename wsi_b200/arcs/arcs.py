"""Arc between two nodes (reference: wsimod/arcs/arcs.py:18-96).  Capacity, preference and
the per-step ``vqip_in`` / ``vqip_out`` totals; the push protocol of arcs.py:112-257 runs in
wsi_b200/csrc/wsi_device.cuh."""
from __future__ import annotations

from ..core import constants
from ..core.core import WSIObj


class Arc(WSIObj):
    def __init__(self, name="", capacity=constants.UNBOUNDED_CAPACITY, preference=1, in_port=None, out_port=None,
                 **kwargs):
        self.name = name
        self.in_port = in_port
        self.out_port = out_port
        self.capacity = capacity
        self.preference = preference
        WSIObj.__init__(self)
        self.__dict__.update(kwargs)
        self.flow_in = 0
        self.flow_out = 0
        self.vqip_in = self.empty_vqip()
        self.vqip_out = self.vqip_in
        in_port.out_arcs[name] = self
        out_port.in_arcs[name] = self
        in_port.out_arcs_type.setdefault(out_port.__class__.__name__, {})[name] = self
        out_port.in_arcs_type.setdefault(in_port.__class__.__name__, {})[name] = self

    def apply_overrides(self, overrides=None):
        overrides = overrides or {}
        self.capacity = overrides.pop("capacity", self.capacity)
        self.preference = overrides.pop("preference", self.preference)
        if overrides:
            print(f"No override behaviour defined for: {overrides.keys()}")


class PullArc(Arc):
    """Only pulls travel along it: pushes and push checks are answered with "nothing" (reference:
    arcs.py:808-850) -- the flattener leaves it out of every push candidate list."""


class PushArc(Arc):
    """Only pushes travel along it (reference: arcs.py:852-893) -- WsiModelDesc.arc_flags tells the
    pull code."""


class SewerArc(Arc):
    """A plain Arc under another name (reference: arcs.py:896-899)."""


class WeirArc(SewerArc):
    """A plain Arc under another name (reference: arcs.py:902-905)."""


SUPPORTED_ARCS = ("Arc", "PullArc", "PushArc", "SewerArc", "WeirArc")
# The arcs that keep a list of individual requests (arcs.py:272-805) are outside the GPU path;
# Model.add_arcs raises for them rather than silently treating them as plain arcs.
UNSUPPORTED_ARCS = ("QueueArc", "AltQueueArc", "DecayArc", "DecayArcAlt")
