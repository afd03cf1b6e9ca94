"""Python wrapper of the C oracle (oracle/wsi_oracle.c).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package wsi_b200
never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

from wsi_b200.backend.capi import ModelBuffers, WsiModelDesc

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libwsi_oracle.so")
_LIB = None


def build(force=False):
    """Compile the C restatement with the system gcc (oracle/Makefile)."""
    inc = os.path.join(os.path.dirname(_HERE), "include")
    deps = [os.path.join(_HERE, "wsi_oracle.c"), os.path.join(inc, "wsi_b200.h"), os.path.join(inc, "wsi_layout.h")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["make", "-C", _HERE, "-B", "-s"])
    return _SO


def lib():
    global _LIB
    if _LIB is None:
        build()  # no-op when the library is newer than its sources
        _LIB = C.CDLL(_SO)
        _LIB.wsi_oracle_desc_size.restype = C.c_int
        if _LIB.wsi_oracle_desc_size() != C.sizeof(WsiModelDesc):
            raise RuntimeError("libwsi_oracle.so was built against another WsiModelDesc layout; make -C oracle -B")
        _LIB.wsi_oracle_run.argtypes = [C.POINTER(WsiModelDesc), C.c_int64, C.c_int64, C.c_int]
        _LIB.wsi_oracle_run.restype = C.c_int
    return _LIB


class OracleRun:
    """Result of one oracle simulation (host numpy arrays)."""

    def __init__(self, buffers):
        self.buffers = buffers
        self.fm = buffers.fm
        self.out_probe = buffers.array("out_probe")
        self.out_mb = buffers.array("out_mb") if buffers.record_mb else None
        self.mb_sums = buffers.array("mb_sums")
        self.mb_count = buffers.array("mb_count")
        self.node_flags = buffers.array("node_flags")
        self.groups = {g: buffers.array("group_" + g) for g in self.fm.groups}


def run(fm, probes=(), t0=0, t1=None, record_mb=False, n_threads=0, rec_t0=None, rec_T=None):
    """Simulate timesteps [t0, t1) of flat model ``fm`` on the CPU oracle."""
    t1 = fm.T if t1 is None else t1
    buf = ModelBuffers(
        fm,
        device="numpy",
        probes=probes,
        rec_t0=t0 if rec_t0 is None else rec_t0,
        rec_T=(t1 - t0) if rec_T is None else rec_T,
        record_mb=record_mb,
    )
    assert buf.desc.n_red == 0 and buf.desc.rec_R == 0  # device-side output reductions are not restated here
    rc = lib().wsi_oracle_run(C.byref(buf.desc), int(t0), int(t1), int(n_threads))
    if rc == 3:
        raise NotImplementedError("the oracle does not restate treatment works / demand / the pull side "
                                  "(their goldens come straight from the reference)")
    if rc != 0:
        raise RuntimeError(f"oracle returned {rc}")
    return OracleRun(buf)
