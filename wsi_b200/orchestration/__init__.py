"""wsi_b200.orchestration: host-side mirror of the reference interface."""
