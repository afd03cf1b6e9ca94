"""Catchment source node (reference: wsimod/nodes/catchment.py:12-88): reads flow and
pollutant concentrations from ``data_input_dict`` and pushes them downstream."""
from .nodes import Node


class Catchment(Node):
    def __init__(self, name, data_input_dict=None):
        super().__init__(name)
        self.data_input_dict = {} if data_input_dict is None else data_input_dict
        self.unrouted_water = self.empty_vqip()
