"""Distribution nodes (reference: wsimod/nodes/distribution.py:14-200).  ``Distribution`` conveys pulls
and denies pushes; with ``leakage`` > 0 every pull through it is enlarged by the leak, which is pushed
to the attached Groundwater (the decorators of distribution.py:14-98).  ``UnlimitedDistribution``
answers every pull in full (its pull handlers replace the decorators, distribution.py:163-168).  The
handlers run on the device (csrc/wsi_works.cuh)."""
from __future__ import annotations

from .nodes import Node


class Distribution(Node):
    def __init__(self, leakage=0, **kwargs):
        self.leakage = leakage
        super().__init__(**kwargs)

    def apply_overrides(self, overrides=None):
        overrides = overrides or {}
        self.leakage = overrides.pop("leakage", self.leakage)
        super().apply_overrides(overrides)


class UnlimitedDistribution(Distribution):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.supplied = self.empty_vqip()
