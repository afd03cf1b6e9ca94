#!/usr/bin/env python
"""Regenerate include/wsi_layout.h from wsi_b200/backend/layout.py."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wsi_b200.backend import layout  # noqa: E402

if __name__ == "__main__":
    path = os.path.join(ROOT, "include", "wsi_layout.h")
    with open(path, "w") as f:
        f.write(layout.c_header())
    print("wrote", path)
