#!/usr/bin/env python
"""Development aid: build variants of libwsi_b200.so (default pollutant set only) with different
compile-time tuning macros, side by side, for A/B runs in one gpurun call.

    python tools/build_variants.py name1:MACRO=V,MACRO2=V2 name2:...
    -> wsi_b200/csrc/variants/libwsi_<name>.so      (select with WSI_B200_LIB=<path>)
"""
import os, sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wsi_b200 import build as wb

vdir = os.path.join(wb.CSRC, "variants")
os.makedirs(vdir, exist_ok=True)


def one(spec):
    name, _, macros = spec.partition(":")
    defs = [m for m in macros.split(",") if m]
    out = os.path.join(vdir, f"libwsi_{name}.so")
    wb.build_extension(force=True, only_default12=True, defines=defs, out=out)
    return out


with ThreadPoolExecutor(max_workers=4) as pool:
    for o in pool.map(one, sys.argv[1:]):
        print(o)
