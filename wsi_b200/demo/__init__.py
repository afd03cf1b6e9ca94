"""wsi_b200.demo: host-side mirror of the reference interface."""
