"""Build tests/hostsim/libwsi_hostsim.so: the kernel logic compiled for the CPU (test only)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libwsi_hostsim.so")
CSRC = os.path.join(os.path.dirname(os.path.dirname(HERE)), "wsi_b200", "csrc")


def build(force=False):
    deps = [os.path.join(HERE, "hostsim.cpp")] + [
        os.path.join(CSRC, f) for f in ("wsi_vq.cuh", "wsi_device.cuh", "wsi_land.cuh", "wsi_nodes.cuh", "wsi_net.cuh", "wsi_works.cuh")
    ] + [os.path.join(os.path.dirname(CSRC), "..", "include", f) for f in ("wsi_b200.h", "wsi_layout.h")]
    if not force and os.path.exists(SO) and all(os.path.getmtime(d) <= os.path.getmtime(SO) for d in deps):
        return SO
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call(
        [gxx, "-std=c++17", "-O1", "-fPIC", "-ffp-contract=off", "-shared", "-o", SO, "hostsim.cpp"], cwd=HERE
    )
    return SO
