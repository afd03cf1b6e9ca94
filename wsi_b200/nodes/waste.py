"""Outlet node (reference: wsimod/nodes/waste.py:10-49): swallows every push."""
from .nodes import Node


class Waste(Node):
    def __init__(self, name):
        super().__init__(name)
