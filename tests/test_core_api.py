"""Host-side VQIP helpers of the drop-in API (wsi_b200/core/core.py) against the reference's own
known answers (reference tests/test_core.py:21-123, hand numbers on the simple pollutant set) --
including the concentration-space variants, which the default path never calls
-- and, where the reference is mounted, against the reference's methods on random VQIPs."""
import numpy as np
import pytest


@pytest.fixture()
def obj():
    from wsi_b200.core import constants
    from wsi_b200.core.core import WSIObj

    constants.set_simple_pollutants()
    yield WSIObj()
    constants.set_default_pollutants()


D1 = {"volume": 10, "phosphate": 0.0001, "temperature": 15}
D2 = {"volume": 5, "phosphate": 0.00005, "temperature": 10}


def test_empty_copy(obj):
    assert obj.empty_vqip() == {"volume": 0, "phosphate": 0, "temperature": 0}
    c = obj.copy_vqip(D1)
    assert c == D1 and c is not D1


def test_blend_and_sum(obj):  # test_core.py:32-44
    assert obj.blend_vqip(D1, D2) == {"volume": 15, "phosphate": 0.00125 / 15, "temperature": 40 / 3}
    assert obj.sum_vqip(D1, D2) == {"volume": 15, "phosphate": 0.0001 + 0.00005, "temperature": 40 / 3}


def test_concentration_conversions(obj):  # test_core.py:46-56
    assert obj.concentration_to_total(D1) == {"volume": 10, "phosphate": 0.001, "temperature": 15}
    assert obj.total_to_concentration(D1) == {"volume": 10, "phosphate": 0.00001, "temperature": 15}


def test_extract_and_extract_c(obj):  # test_core.py:58-74
    assert obj.extract_vqip(D1, D2) == {"volume": 5, "phosphate": 0.0001 - 0.00005, "temperature": 15}
    assert obj.extract_vqip_c(D1, D2) == {"volume": 5, "phosphate": (0.0001 * 10 - 0.00005 * 5) / 5, "temperature": 15}


def test_distill_and_vchange(obj):  # test_core.py:76-102
    assert obj.v_distill_vqip(D1, 5) == {"volume": 5, "phosphate": 0.0001, "temperature": 15}
    assert obj.v_distill_vqip_c(D1, 5) == {"volume": 5, "phosphate": 0.0002, "temperature": 15}
    assert obj.v_change_vqip(D1, 5) == {"volume": 5, "phosphate": 0.00005, "temperature": 15}
    assert obj.v_change_vqip_c(D1, 5) == {"volume": 5, "phosphate": 0.0001, "temperature": 15}


def test_ds_and_compare(obj):  # test_core.py:104-123
    assert obj.ds_vqip(D1, D2) == {"volume": 5, "phosphate": 0.00005, "temperature": 0}
    assert obj.ds_vqip_c(D1, D2) == {"volume": 5, "phosphate": 0.00075, "temperature": 0}
    assert not obj.compare_vqip(D1, D2)
    assert obj.compare_vqip(D1, D1)


@pytest.mark.reference
def test_every_helper_equals_the_reference_method_on_random_vqips():
    """Same method names, same results bit for bit, default 12-pollutant set."""
    import sys

    sys.path.append("/root/reference")  # appended: its own `tests` package must not shadow this one
    sys.dont_write_bytecode = True
    try:
        from wsimod.core import constants as rc
        from wsimod.core.core import WSIObj as Ref
    finally:
        sys.path.remove("/root/reference")

    from wsi_b200.core import constants as mc
    from wsi_b200.core.core import WSIObj as Mine

    rc.set_default_pollutants()
    mc.set_default_pollutants()
    ref, mine = Ref(), Mine()
    rng = np.random.default_rng(7)

    def vq():
        d = {p: float(rng.uniform(0, 3)) for p in rc.POLLUTANTS}
        d["volume"] = float(rng.uniform(1, 50))
        return d

    two = ["blend_vqip", "sum_vqip", "extract_vqip", "extract_vqip_c", "ds_vqip", "ds_vqip_c", "compare_vqip"]
    one = ["concentration_to_total", "total_to_concentration", "copy_vqip"]
    withv = ["v_distill_vqip", "v_distill_vqip_c", "v_change_vqip", "v_change_vqip_c"]
    for _ in range(25):
        a, b = vq(), vq()
        b["volume"] = a["volume"] * float(rng.uniform(0.1, 0.9))  # extractable
        for name in two:
            assert getattr(mine, name)(dict(a), dict(b)) == getattr(ref, name)(dict(a), dict(b)), name
        for name in one:
            assert getattr(mine, name)(dict(a)) == getattr(ref, name)(dict(a)), name
        v = float(rng.uniform(0.1, 0.9)) * a["volume"]
        for name in withv:
            assert getattr(mine, name)(dict(a), v) == getattr(ref, name)(dict(a), v), name
    assert mine.empty_vqip() == ref.empty_vqip()
