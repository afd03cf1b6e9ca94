"""Known answers of the reference's own unit tests, per device function (CPU, host build of the kernel
headers: tests/hostsim).  Configuration as in those tests: VQIP = {volume, phosphate (additive),
temperature (non-additive)}.  Each case cites the reference test it restates."""
import ctypes as C

import numpy as np
import pytest


@pytest.fixture(scope="module")
def lib():
    from hostsim.build import build

    h = C.CDLL(build())
    dp = C.POINTER(C.c_double)
    h.wsi_hostsim_vq.argtypes = [C.c_int, dp, dp, C.c_double, dp]
    h.wsi_hostsim_tank.argtypes = [C.c_int, C.c_double, dp, dp, C.c_double, dp]
    h.wsi_hostsim_decay.argtypes = [dp, C.c_int, C.c_double, C.c_double, C.c_double, dp]
    for f in (h.wsi_hostsim_vq, h.wsi_hostsim_tank, h.wsi_hostsim_decay):
        f.restype = None
    return h


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def vq(d):
    return np.array([d["volume"], d["phosphate"], d["temperature"]], dtype=np.float64)


def vq_op(lib, op, a, b=None, s=0.0):
    a = vq(a)
    b = vq(b) if b is not None else np.zeros(3)
    out = np.zeros(3)
    lib.wsi_hostsim_vq(op, _p(a), _p(b), float(s), _p(out))
    return out


class Tank:
    def __init__(self, lib, capacity, storage):
        self.lib, self.capacity, self.storage = lib, float(capacity), vq(storage)

    def call(self, op, v=None, s=0.0):
        v = vq(v) if v is not None else np.zeros(3)
        out = np.zeros(3)
        self.lib.wsi_hostsim_tank(op, self.capacity, _p(self.storage), _p(v), float(s), _p(out))
        return out


D1 = {"volume": 10, "phosphate": 0.0001, "temperature": 15}
D2 = {"volume": 5, "phosphate": 0.00005, "temperature": 10}


def test_sum_vqip(lib):  # tests/test_core.py:39-44
    assert np.array_equal(vq_op(lib, 0, D1, D2), [15, 0.0001 + 0.00005, 40 / 3])


def test_extract_vqip(lib):  # tests/test_core.py:58-63
    assert np.array_equal(vq_op(lib, 1, D1, D2), [5, 0.0001 - 0.00005, 15])


def test_v_change_vqip(lib):  # tests/test_core.py:88-93
    assert np.array_equal(vq_op(lib, 2, D1, s=5), [5, 0.00005, 15])


def test_compare_vqip(lib):  # tests/test_core.py:116-121
    assert vq_op(lib, 3, D1, D2)[0] == 0.0
    assert vq_op(lib, 3, D1, D1)[0] == 1.0


def test_empty_vqip_is_recognised(lib):  # core.py:32-40; the empty-request fast path relies on it
    assert vq_op(lib, 4, {"volume": 0, "phosphate": 0, "temperature": 17})[0] == 1.0
    assert vq_op(lib, 4, {"volume": 0, "phosphate": 1e-300, "temperature": 0})[0] == 0.0


def test_tank_ds(lib):  # tests/test_tanks.py:35-49: push 2 into 5 of 10, ds = storage - storage_
    t = Tank(lib, 10, {"volume": 5, "phosphate": 0.4, "temperature": 10})
    before = t.storage.copy()
    t.call(0, {"volume": 2, "phosphate": 0.01, "temperature": 15})
    diff = t.storage - before
    assert np.allclose(diff[:2], [2, 0.01], rtol=0, atol=1e-16)


def test_pull_ponded(lib):  # tests/test_tanks.py:51-58
    t = Tank(lib, 10, {"volume": 15, "phosphate": 0.4, "temperature": 10})
    assert np.allclose(t.call(3), [5, 0.4 / 3, 10], rtol=0, atol=1e-15)
    assert t.storage[0] == 10


def test_get_excess(lib):  # tests/test_tanks.py:74-84 (the kernels only ever read the volume of a check)
    t = Tank(lib, 10, {"volume": 7.5, "phosphate": 0.4, "temperature": 10})
    assert t.call(4, s=-1)[0] == 2.5
    assert t.call(4, s=1)[0] == pytest.approx(1, abs=1e-15)


def test_push_storage(lib):  # tests/test_tanks.py:86-98
    t = Tank(lib, 10, {"volume": 7.5, "phosphate": 0.4, "temperature": 10})
    d2 = {"volume": 5, "phosphate": 0.4, "temperature": 15}
    assert np.allclose(t.call(0, d2), [2.5, 0.2, 15], rtol=0, atol=1e-15)
    assert np.array_equal(t.call(1, d2), [0, 0, 0])
    assert t.storage[0] == 15  # forced in above the capacity


def test_pull_storage(lib):  # tests/test_tanks.py:100-112
    t = Tank(lib, 10, {"volume": 7.5, "phosphate": 0.4, "temperature": 10})
    assert np.allclose(t.call(2, s=5), [5, 0.4 * 5 / 7.5, 10], rtol=0, atol=1e-15)
    assert np.allclose(t.call(2, s=5), [2.5, 0.4 * 2.5 / 7.5, 10], rtol=0, atol=1e-15)
    assert np.array_equal(t.call(2, s=5), [0, 0, 0])  # an empty tank returns the empty VQIP (tanks.py:227-228)


def test_pull_pollutants(lib):  # tests/test_tanks.py:114-125
    t = Tank(lib, 10, {"volume": 7.5, "phosphate": 0.4, "temperature": 10})
    d2 = {"volume": 5, "phosphate": 0.1, "temperature": 10}
    assert np.allclose(t.call(5, d2), [5, 0.1, 10], rtol=0, atol=1e-15)
    assert np.allclose(t.call(5, d2), [2.5, 0.1, 10], rtol=0, atol=1e-15)


def test_evaporate(lib):  # tests/test_tanks.py:143-151
    t = Tank(lib, 10, {"volume": 7.5, "phosphate": 0.4, "temperature": 10})
    assert t.call(7, s=10)[0] == 7.5
    assert np.array_equal(t.storage, [0, 0.4, 10])


def test_residence_tank_pull_outflow(lib):  # tests/test_tanks.py:153-159
    t = Tank(lib, 0, {"volume": 7.5, "phosphate": 0.4, "temperature": 10})
    assert np.allclose(t.call(6, s=3), [2.5, 0.4 / 3, 10], rtol=0, atol=1e-15)


def test_decay_tank(lib):  # tests/test_tanks.py:161-188
    t = Tank(lib, 10, {"volume": 8, "phosphate": 0.4, "temperature": 10})
    t.call(2, s=2)
    assert np.allclose(t.storage - [8, 0.4, 10], [-2, -0.1, 0], rtol=0, atol=1e-16)  # decay_ds before any decay
    out = np.zeros(3)
    lib.wsi_hostsim_decay(_p(t.storage), 1, 0.001, 1.005, 15.0, _p(out))
    decayed = 0.3 * 0.001 * 1.005 ** (15 - 20)
    assert np.allclose(t.storage, [6, 0.3 - decayed, 10], rtol=0, atol=1e-16)
    assert out[0] == pytest.approx(decayed, abs=1e-16)
    assert out[1] == pytest.approx(0.3, abs=1e-16)  # storage_ keeps the value before the decay
