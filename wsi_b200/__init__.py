"""wsi_b200: B200-native backend for WSIMOD's land / tank / arc hot path behind the reference's own
``Model.load`` / ``Model.run`` API.  The CUDA library lives in ``wsi_b200/csrc`` and is loaded
through ``wsi_b200.backend.capi``; there is no CPU execution path."""
