"""Known answers of the reference's land tests, surface class by surface class (CPU: host build of the
kernel headers; ``-m gpu``: the CUDA library through the C ABI).  The reference calls one process of one surface (tests/test_land.py); here a Land node
holding just that surface -- no deposition data, no pollutant load, so that process is the only one that
acts -- is built through this package's Model API, flattened and stepped once, and the surface's own
state and flow diagnostics are compared with the reference's expected values."""
import ctypes as C
from math import exp

import numpy as np
import pandas as pd
import pytest


@pytest.fixture(scope="module", params=["host build", pytest.param("gpu", marks=pytest.mark.gpu)])
def hostsim(request):
    """Where the Land is stepped: the host build of the kernel headers (CPU suite), or libwsi_b200.so through the
    C ABI on cuda:0 (``-m gpu``) -- the same assertions either way."""
    if request.param == "gpu":
        return None
    from hostsim.build import build
    from wsi_b200.backend.capi import WsiModelDesc

    lib = C.CDLL(build())
    lib.wsi_hostsim_run.argtypes = [C.POINTER(WsiModelDesc), C.c_int64, C.c_int64, C.c_int]
    lib.wsi_hostsim_run.restype = C.c_int
    return lib


def _step(hostsim, surface, forcing_rows):
    """Run len(forcing_rows) timesteps of Land{surface}; returns per step (storage VQIP, SURF_OUT volumes)."""
    from wsi_b200.backend.capi import ModelBuffers
    from wsi_b200.backend.probes import SURF_OUT, ProbeSet
    from wsi_b200.core import constants
    from wsi_b200.orchestration.model import Model

    constants.set_simple_pollutants()  # phosphate (additive), temperature
    constants.FLOAT_ACCURACY = 1e-6
    try:
        dates = [pd.Timestamp("2000-01-01") + pd.Timedelta(days=i) for i in range(len(forcing_rows))]
        inputs = {}
        for d, (p, e, t) in zip(dates, forcing_rows):
            inputs.update({("precipitation", d): p, ("et0", d): e, ("temperature", d): t})
        m = Model()
        m.add_nodes([{"type_": "Land", "name": "land", "data_input_dict": inputs, "surfaces": [dict(surface, surface="s")]}])
        m.dates = dates
        fm = m.flatten()
        ps = ProbeSet(fm).add_standard()
        if hostsim is None:
            from wsi_b200.backend.engine import GpuRunner

            run = GpuRunner(fm, probes=ps.probes, rec_t0=0, rec_T=fm.T, record_mb=False, device="cuda:0")
            run.run(0, fm.T)
            run.synchronize()
            blocks = ps.split(run.out_probe())
        else:
            buf = ModelBuffers(fm, device="numpy", probes=ps.probes, rec_t0=0, rec_T=fm.T, record_mb=False, engine="replica")
            assert hostsim.wsi_hostsim_run(C.byref(buf.desc), 0, fm.T, 0) == 0
            blocks = ps.split(buf.array("out_probe"))
        order = [fm.pollutants.index(p) + 1 for p in ("phosphate", "temperature")]
        storage = blocks["surf_storage"][:, 0, :, 0][:, [0] + order]  # volume, phosphate, temperature
        out = {k.replace("out_", ""): blocks["surf_out"][:, 0, i, 0] for i, k in enumerate(SURF_OUT)}
        return storage, out
    finally:
        constants.set_default_pollutants()
        constants.FLOAT_ACCURACY = 1e-11


PERVIOUS = dict(type_="PerviousSurface", depth=0.5, area=1.5, total_porosity=0.45, field_capacity=0.35, wilting_point=0.12,
                infiltration_capacity=0.4, surface_coefficient=0.04, percolation_coefficient=0.6, et0_coefficient=0.4,
                ihacres_p=12)


def _soil_temperature(tank_temperature, air_temperature):
    """PerviousSurface.run also applies calculate_soil_temperature (land.py:1060-1079, weights land.py:809-812)
    after ihacres: the temperature the reference's ihacres test expects is its input, tests/test_land.py:482."""
    return (tank_temperature * 0.1 + air_temperature * 0.6 + 10 * 0.1) / (0.6 + 0.1 + 0.1)


def test_impervious_precipitation_evaporation(hostsim):  # tests/test_land.py:174-212
    surf = dict(type_="ImperviousSurface", area=1.5, et0_to_e=0.9, pore_depth=0.015)
    storage, out = _step(hostsim, surf, [(0.1, 0.01, 10), (0.0, 0.02, 15)])
    assert out["precipitation"][0] == pytest.approx(0.1 * 1.5, abs=1e-17)
    assert out["evaporation"][0] == pytest.approx(0.01 * 1.5 * 0.9, abs=1e-17)
    assert out["precipitation"][1] == 0.0
    assert out["evaporation"][1] == pytest.approx(0.02 * 1.5 * 0.9, abs=1e-17)
    # the reference's test stops before push_to_sewers (land.py:699-721); with no sewer attached the ponded
    # water is pushed nowhere and forced back into the tank, so the storage it expects is what remains
    assert storage[0, 0] == pytest.approx((0.1 - 0.01 * 0.9) * 1.5, abs=1e-16)
    assert storage[0, 2] == 10
    assert storage[1, 0] == pytest.approx((0.1 - (0.01 + 0.02) * 0.9) * 1.5, abs=1e-16)


def test_ihacres_above_field_capacity(hostsim):  # tests/test_land.py:254-323
    storage, out = _step(hostsim, dict(PERVIOUS, initial_storage=0.4 * 0.5 * 1.5), [(0.1, 0.01, 10)])
    outflow_ = 0.1 * (1 - ((0.45 * 0.5 - 0.4 * 0.5) / (0.45 * 0.5 - 0.35 * 0.5)) ** 12)
    through = (0.1 - 0.01 * 0.4 - outflow_ * 0.04) * 1.5
    temperature = (0 * 0.4 * 0.5 * 1.5) + (through * 10) / (through + 0.4 * 0.5 * 1.5)
    assert out["precipitation"][0] == pytest.approx(0.1 * 1.5, abs=1e-16)
    assert out["evaporation"][0] == pytest.approx(0.01 * 1.5 * 0.4, abs=1e-16)
    assert out["percolation"][0] == pytest.approx(outflow_ * (1 - 0.04) * 0.6 * 1.5, abs=1e-15)
    assert out["subsurface"][0] == pytest.approx(outflow_ * (1 - 0.04) * (1 - 0.6) * 1.5, abs=1e-15)
    assert out["surface"][0] == pytest.approx(outflow_ * 0.04 * 1.5, abs=1e-15)
    assert storage[0, 0] == pytest.approx((0.1 - outflow_ - 0.01 * 0.4 + 0.4 * 0.5) * 1.5, abs=1e-15)
    assert storage[0, 2] == pytest.approx(_soil_temperature(temperature, 10), abs=1e-14)


def test_ihacres_below_wilting_point(hostsim):  # tests/test_land.py:325-380
    storage, out = _step(hostsim, dict(PERVIOUS, initial_storage=0.11 * 0.5 * 1.5), [(0.0, 0.01, 10)])
    evaporation_ = 0.01 * 0.4 * exp(2 * (1 - (0.45 * 0.5 - 0.11 * 0.5) / (0.45 * 0.5 - 0.12 * 0.5))) * 1.5
    assert out["precipitation"][0] == 0.0
    assert out["evaporation"][0] == pytest.approx(evaporation_, abs=1e-16)
    assert out["percolation"][0] == 0.0 and out["subsurface"][0] == 0.0 and out["surface"][0] == 0.0
    assert storage[0, 0] == pytest.approx(0.11 * 0.5 * 1.5 - evaporation_, abs=1e-15)


def test_ihacres_infiltration_excess(hostsim):  # tests/test_land.py:382-455
    surf = dict(PERVIOUS, infiltration_capacity=0.05, initial_storage=0.4 * 0.5 * 1.5)
    storage, out = _step(hostsim, surf, [(0.1, 0.01, 10)])
    outflow_ = 0.05 * (1 - ((0.45 * 0.5 - 0.4 * 0.5) / (0.45 * 0.5 - 0.35 * 0.5)) ** 12)
    through = (0.05 - 0.01 * 0.4 - outflow_ * 0.04) * 1.5
    temperature = (0 * 0.4 * 0.5 * 1.5) + (through * 10) / (through + 0.4 * 0.5 * 1.5)
    assert out["percolation"][0] == pytest.approx(outflow_ * (1 - 0.04) * 0.6 * 1.5, abs=1e-15)
    assert out["subsurface"][0] == pytest.approx(outflow_ * (1 - 0.04) * (1 - 0.6) * 1.5, abs=1e-15)
    assert out["surface"][0] == pytest.approx(outflow_ * 0.04 * 1.5 + 0.05 * 1.5, abs=1e-15)
    assert storage[0, 0] == pytest.approx((0.05 - outflow_ - 0.01 * 0.4 + 0.4 * 0.5) * 1.5, abs=1e-15)
    assert storage[0, 2] == pytest.approx(_soil_temperature(temperature, 10), abs=1e-14)
