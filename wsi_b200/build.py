"""Build the CUDA extension in-tree (wsi_b200/csrc/libwsi_b200.so) for sm_100a."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
SO = os.path.join(CSRC, "libwsi_b200.so")
SOURCES = ["wsi_api.cu"]
HEADERS = [
    os.path.join(CSRC, "wsi_vq.cuh"),
    os.path.join(CSRC, "wsi_device.cuh"),
    os.path.join(CSRC, "wsi_land.cuh"),
    os.path.join(CSRC, "wsi_nodes.cuh"),
    os.path.join(ROOT, "include", "wsi_b200.h"),
    os.path.join(ROOT, "include", "wsi_layout.h"),
]

NVCC_FLAGS = [
    "-std=c++17",
    "-O3",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    # the reference rounds every multiply and add separately (CPython floats); keep the
    # kernels' own arithmetic contraction-free so results track it (libdevice pow/exp/log
    # keep their explicit fma)
    "-fmad=false",
    "-Xcompiler", "-fPIC",
    "-shared",
]


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found")
    return p


def is_current():
    if not os.path.exists(SO):
        return False
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS
    return all(os.path.getmtime(d) <= t for d in deps)


def build_extension(force=False, verbose=False):
    """Compile libwsi_b200.so (cross-compiles without a GPU)."""
    if not force and is_current():
        return SO
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO] + SOURCES
    subprocess.check_call(cmd, cwd=CSRC)
    return SO


if __name__ == "__main__":
    print(build_extension(force=True, verbose=True))
