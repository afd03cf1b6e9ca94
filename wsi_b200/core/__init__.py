"""wsi_b200.core: host-side mirror of the reference interface."""
