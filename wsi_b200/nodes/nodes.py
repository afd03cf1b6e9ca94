"""``Node`` base class and the node registry (reference: wsimod/nodes/nodes.py:15-80, 759).

A node here is a *description*: name, parameters, initial state and its arcs.  The
push/pull message passing of the reference (nodes.py:148-478) is executed by the device
kernels; calling the per-node orchestration methods directly is not supported."""
from __future__ import annotations

import logging

from ..core.core import WSIObj

NODES_REGISTRY: dict = {}


class Node(WSIObj):
    def __init_subclass__(cls, **kwargs):
        super().__init_subclass__(**kwargs)
        if cls.__name__ in NODES_REGISTRY:
            logging.warning(f"Overwriting {cls.__name__} in NODES_REGISTRY with {cls}")
        NODES_REGISTRY[cls.__name__] = cls

    def __init__(self, name, data_input_dict=None):
        self.in_arcs = {}
        self.out_arcs = {}
        self.in_arcs_type = {x: {} for x in NODES_REGISTRY}
        self.out_arcs_type = {x: {} for x in NODES_REGISTRY}
        self.name = name
        self.t = None
        self.data_input_dict = data_input_dict
        super().__init__()

    def apply_overrides(self, overrides=None):
        overrides = overrides or {}
        content = overrides.pop("filename", None)
        if isinstance(content, str):
            from ..orchestration.model import read_csv

            self.data_input_dict = read_csv(content)
        elif content:
            raise RuntimeError("Not recognised format for data_input_dict")
        if overrides:
            print(f"No override behaviour defined for: {overrides.keys()}")

    def get_data_input(self, var):
        return self.data_input_dict[(var, self.t)]

    def _device_only(self, what):
        raise NotImplementedError(
            f"{type(self).__name__}.{what} runs inside the CUDA kernels of wsi_b200; "
            "call Model.run instead of stepping nodes by hand"
        )

    def push_distributed(self, *a, **k):
        self._device_only("push_distributed")

    def pull_distributed(self, *a, **k):
        self._device_only("pull_distributed")

    def end_timestep(self):
        pass

    def reinit(self):
        pass


NODES_REGISTRY[Node.__name__] = Node
