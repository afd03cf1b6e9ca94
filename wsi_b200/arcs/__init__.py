"""wsi_b200.arcs: host-side mirror of the reference interface."""
