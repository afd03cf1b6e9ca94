"""The kernel logic (wsi_b200/csrc/*.cuh) compiled for the CPU and checked against the golden
vectors, so the GPU-less CI still exercises the code the CUDA kernels are made of.  The host
build is test infrastructure (tests/hostsim); the product never loads it."""
import ctypes as C

import numpy as np
import pytest

from parity_util import compare_blocks, compare_mb, load_golden

GOLDENS = ["quickstart", "quickstart_simple", "quickstart_p5", "quickstart_p14", "fixture3607", "config2", "rivernet", "rivertree", "config3", "hourly",
           "oxford", "fixture3607_full", "city", "queuegw", "reservoirs", "irrigation", "distribution"]


@pytest.fixture(scope="module")
def hostsim():
    from hostsim.build import build
    from wsi_b200.backend.capi import WsiModelDesc

    lib = C.CDLL(build())
    lib.wsi_hostsim_run.argtypes = [C.POINTER(WsiModelDesc), C.c_int64, C.c_int64, C.c_int]
    lib.wsi_hostsim_run.restype = C.c_int
    return lib


def _run(lib, fm, probes, force_runtime=0, record_mb=True, engine="replica"):
    from wsi_b200.backend.capi import ModelBuffers

    buf = ModelBuffers(fm, device="numpy", probes=probes, rec_t0=0, rec_T=fm.T, record_mb=record_mb, engine=engine)
    assert lib.wsi_hostsim_run(C.byref(buf.desc), 0, fm.T, force_runtime) == 0
    return buf


@pytest.mark.parametrize("name", GOLDENS)
def test_kernel_logic_matches_reference_outputs(hostsim, name):
    from wsi_b200.backend.probes import ProbeSet

    fm, ref = load_golden(name)
    ps = ProbeSet(fm).add_standard()
    buf = _run(hostsim, fm, ps.probes)
    ok, reports = compare_blocks(ps.split(buf.array("out_probe")), ref, fm)
    ok2, rep = compare_mb(buf.array("out_mb"), ref["mb"], fm)
    assert ok and ok2, "\n".join(reports + [rep])


def test_runtime_pollutant_config_equals_static_config(hostsim):
    """CfgRuntime<12> (pollutant planes read from the descriptor) and CfgDefault12 (compile-time
    planes) are the same arithmetic."""
    from wsi_b200.backend.probes import ProbeSet

    fm, _ = load_golden("config2")
    ps = ProbeSet(fm).add_standard()
    a = _run(hostsim, fm, ps.probes, 0, record_mb=False)
    b = _run(hostsim, fm, ps.probes, 1, record_mb=False)
    assert np.array_equal(a.array("out_probe"), b.array("out_probe"))


def test_kernel_logic_equals_oracle_on_ensemble(hostsim):
    """config-2 ensemble (the benchmark workload) at a CPU-sized member count."""
    from oracle import oracle
    from wsi_b200 import ensemble as E
    from wsi_b200.backend.probes import ProbeSet

    fm = E.build_config2(R=16, T=400)
    ps = ProbeSet(fm).add_flows()
    buf = _run(hostsim, fm, ps.probes, record_mb=False)
    ora = oracle.run(fm, probes=ps.probes)
    assert np.array_equal(buf.array("out_probe"), ora.out_probe)
    np.testing.assert_allclose(buf.array("mb_sums"), ora.mb_sums, rtol=1e-12, atol=1e-9)
    assert np.array_equal(buf.array("mb_count"), ora.mb_count)


def test_network_engine_matches_reference_outputs(hostsim):
    """Level-order routing (wsi_net.cuh) reproduces the reference's sequential, upstream-first
    river routing bit for bit on a dendritic network."""
    from wsi_b200.backend.probes import ProbeSet

    fm, ref = load_golden("rivertree")
    ps = ProbeSet(fm).add_standard()
    buf = _run(hostsim, fm, ps.probes, engine="network")
    assert buf.engine == 1 and buf.net["n_levels"] >= 4
    ok, reports = compare_blocks(ps.split(buf.array("out_probe")), ref, fm)
    ok2, rep = compare_mb(buf.array("out_mb"), ref["mb"], fm)
    assert ok and ok2, "\n".join(reports + [rep])
    seq = _run(hostsim, fm, ps.probes, engine="replica")
    assert np.array_equal(buf.array("out_probe"), seq.array("out_probe"))
    assert np.array_equal(buf.array("out_mb"), seq.array("out_mb"))


def test_network_engine_rejects_non_dendritic_models():
    from wsi_b200.backend import network
    from wsi_b200.backend.capi import ModelBuffers

    fm, _ = load_golden("rivernet")  # has Land, Sewer, a preference split ...
    assert network.plan_network(fm) is None and network.why_not(fm)
    with pytest.raises(ValueError):
        ModelBuffers(fm, device="numpy", engine="network")


def test_reduced_outputs_and_member_window(hostsim):
    """Objectives on the device: a reduced probe keeps the running sum and maximum of a plane per
    member (model.py:1177-1197 needs no more of the series), and the series themselves are kept for
    a window of members only -- both must equal what the full record gives."""
    from wsi_b200 import ensemble as E
    from wsi_b200.backend.capi import ModelBuffers
    from wsi_b200.backend.probes import ProbeSet

    fm = E.build_config2(R=24, T=60)
    outlet = [fm.arc_names.index("river-outlet")]
    ps = ProbeSet(fm).add_flows(outlet)
    full = _run(hostsim, fm, ps.probes, record_mb=False).array("out_probe")  # [T, V, R]
    buf = ModelBuffers(fm, device="numpy", probes=ps.probes, rec_t0=0, rec_T=fm.T, reduce_probes=ps.probes,
                       rec_members=(5, 7), engine="replica")
    assert hostsim.wsi_hostsim_run(C.byref(buf.desc), 0, fm.T, 0) == 0
    assert np.array_equal(buf.array("out_probe"), full[:, :, 5:12])
    red = buf.array("red_out")
    assert np.array_equal(red[1], full.max(axis=0))
    want = np.zeros_like(red[0])
    for t in range(fm.T):  # the same left-to-right additions
        want = want + full[t]
    assert np.array_equal(red[0], want)
