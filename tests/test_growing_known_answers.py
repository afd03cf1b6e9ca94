"""Known answers of the reference's GrowingSurface tests, one process at a time (CPU, host build of the kernel
headers).  The reference prepares a surface (tests/test_land.py:516-562 ``create_growing_surface``) and calls ONE
of its methods; here the same surface is built through this package's Model API as the only surface of a Land,
``surface_phase`` (csrc/wsi_land.cuh) is stepped once with only that process (and what the reference's test
calls before it) switched on -- the host build's WSI_PROC mask; every CUDA build compiles the switch away -- and
the surface's storage, nutrient pools, outgoing flows and mass-balance shares are compared with the values the
reference's tests expect."""
import ctypes as C

import numpy as np
import pandas as pd
import pytest

PROC = {n: i for i, n in enumerate(
    ["crop_cover", "dry_deposition", "wet_deposition", "simple_deposition", "ihacres", "flushing", "fertiliser", "manure",
     "soil_temperature", "temperature_factor", "moisture_factor", "soil_pool", "crop_uptake", "erosion", "denitrification",
     "adsorption"])}

IVOL = {"phosphate": 11, "nitrate": 2.5, "nitrite": 1.5, "ammonia": 0.1, "org-nitrogen": 0.2, "org-phosphorus": 3,
        "volume": 0.32, "temperature": 11}
ISOIL = {"phosphate": 1.2, "ammonia": 0.2, "nitrate": 0.3, "nitrite": 0.4, "org-nitrogen": 2, "org-phosphorus": 4}


@pytest.fixture(scope="module")
def hostsim():
    from hostsim.build import build
    from wsi_b200.backend.capi import WsiModelDesc

    lib = C.CDLL(build())
    lib.wsi_hostsim_run.argtypes = [C.POINTER(WsiModelDesc), C.c_int64, C.c_int64, C.c_int]
    lib.wsi_hostsim_run.restype = C.c_int
    lib.wsi_hostsim_process_mask.argtypes = [C.c_uint]
    lib.wsi_hostsim_process_mask.restype = None
    return lib


class Result:
    """State of the surface after the step, by the reference's names."""

    def __init__(self, fm, groups):
        self.fm, self.g = fm, groups

    def vq(self, field):
        from wsi_b200.backend import layout as L

        o = L.SURF.offset(field, self.fm.V)
        col = self.g["surf"][o:o + self.fm.V, 0]
        d = {"volume": col[0]}
        d.update({p: col[i + 1] for i, p in enumerate(self.fm.pollutants)})
        return d

    def pool(self, name):
        from wsi_b200.backend import layout as L

        o = L.SURF.offset("pool_" + name, self.fm.V)
        return {"N": self.g["surf"][o, 0], "P": self.g["surf"][o + 1, 0]}


def run_process(hostsim, processes, date="2000-05-01", forcing=(0.0, 0.0, 0.0), prepare=None, surface_data=None):
    """One timestep of Land{the reference tests' growing surface} with only ``processes`` switched on."""
    from wsi_b200.backend.capi import ModelBuffers
    from wsi_b200.core import constants
    from wsi_b200.orchestration.model import Model

    constants.set_default_pollutants()
    d = pd.Timestamp(date)
    inputs = {("precipitation", d): forcing[0], ("et0", d): forcing[1], ("temperature", d): forcing[2]}
    surf = dict(type_="GrowingSurface", surface="s", rooting_depth=0.5, area=1.5, initial_storage=dict(IVOL),
                initial_soil_storage=dict(ISOIL), crop_factor_stages=[0.0, 0.0, 0.3, 0.3, 1.2, 1.2, 0.325, 0.0, 0.0],
                crop_factor_stage_dates=[0, 90, 91, 121, 161, 213, 244, 245, 366], sowing_day=91, harvest_day=244,
                ET_depletion_factor=0.55, wilting_point=0.1, field_capacity=0.2)
    # fertiliser() and manure() are entries of every GrowingSurface's inflows (land.py:1279-1280) and read these
    # monthly inputs, so a model that steps needs them; zero unless the test says otherwise
    data = {(v + k, d.to_period("M")): 0.0 for v in ("nhx-", "noy-", "srp-") for k in ("fertiliser", "manure")}
    data.update(surface_data or {})
    surf["data_input_dict"] = data
    m = Model()
    m.add_nodes([{"type_": "Land", "name": "land", "data_input_dict": inputs, "surfaces": [surf]}])
    m.dates = [d]
    s = m.nodes["land"].surfaces[0]
    if prepare:
        prepare(s)
    fm = m.flatten()
    buf = ModelBuffers(fm, device="numpy", probes=(), rec_t0=0, rec_T=1, record_mb=False, engine="replica")
    mask = 0
    for p in processes:
        mask |= 1 << PROC[p]
    hostsim.wsi_hostsim_process_mask(mask)
    try:
        assert hostsim.wsi_hostsim_run(C.byref(buf.desc), 0, 1, 0) == 0
    finally:
        hostsim.wsi_hostsim_process_mask(0xFFFFFFFF)
    return Result(fm, buf.groups_numpy())


def only(d, **nonzero):
    """An otherwise empty VQIP with the given additive entries (the reference's ``empty_vqip`` + assignments)."""
    for k, v in d.items():
        if k in ("temperature", "do", "ph"):
            continue  # non-additive lanes of a mass-balance share are never formed (land.py:441-449)
        assert v == pytest.approx(nonzero.get(k.replace("-", "_"), 0.0), abs=1e-15, rel=1e-13), k
    return True


def test_initial_pools(hostsim):  # tests/test_land.py:564-590 (GrowingSurface.__init__ land.py:1293-1330)
    r = run_process(hostsim, [])
    assert r.pool("dis_inorg") == {"N": pytest.approx(2.5 + 1.5 + 0.1), "P": 11}
    assert r.pool("dis_org") == {"N": 0.2, "P": 3}
    assert r.pool("ads_inorg") == {"N": pytest.approx(0.3 + 0.4 + 0.2, abs=1e-15), "P": 1.2}
    assert r.pool("fast") == {"N": 2, "P": 4}
    assert r.vq("storage")["volume"] == 0.32 and r.vq("storage")["nitrate"] == 2.5  # nothing ran, nothing moved


def test_crop_uptake(hostsim):  # tests/test_land.py:912-941 (calc_crop_cover, then calc_crop_uptake land.py:1858-1926)
    def prepare(s):
        s.days_after_sow = pd.Timestamp("2000-05-01").dayofyear - s.sowing_day

    r = run_process(hostsim, ["crop_cover", "crop_uptake"], prepare=prepare)
    assert only(r.vq("mb_in")) and only(r.vq("mb_out"), nitrate=4.7281041425055e-05, phosphate=6.5668113090354e-06)
    st = r.vq("storage")
    for k, v in IVOL.items():
        take = {"nitrate": 4.7281041425055e-05, "phosphate": 6.5668113090354e-06}.get(k, 0.0)
        assert st[k] == pytest.approx(v - take, abs=1e-14), k
    di = r.pool("dis_inorg")
    assert di["N"] == pytest.approx(2.5 + 0.1 + 1.5 - 4.7281041425055e-05, abs=1e-14)
    assert di["P"] == pytest.approx(11 - 6.5668113090354e-06, abs=1e-14)


@pytest.mark.parametrize("rain, phosphate, solids", [
    (1.5, 0.0019497135091285024, 1.0560948174446056),  # tests/test_land.py:943-964 high rain
    (0.1, 1.0244298083948e-06, 0.0005548994795471758),  # :966-987 medium rain
    (0.005, 8.3285749289e-09, 1.7605117735786e-06),  # :989-1010 low rain
])
def test_erosion(hostsim, rain, phosphate, solids):  # ihacres, then erosion land.py:1928-2094
    r = run_process(hostsim, ["ihacres", "erosion"], forcing=(rain, 0.01, 10))
    mb_in = r.vq("mb_in")
    assert mb_in["phosphate"] == pytest.approx(phosphate, rel=1e-9)  # the reference asserts 7 decimal places
    assert mb_in["solids"] == pytest.approx(solids, rel=1e-9)
    # what erosion mobilised is what the three outgoing flows carry beyond ihacres' own (solids: only erosion's)
    flows = [r.vq(f) for f in ("flow_a", "flow_b", "flow_c")]
    assert sum(f["solids"] for f in flows) == pytest.approx(solids, rel=1e-12)
    # erode_P takes adsorbed inorganic P and humus P in proportion (nutrient_pool.py:367-394); humus starts empty
    assert r.pool("ads_inorg")["P"] == pytest.approx(1.2 - phosphate, rel=1e-12)


def test_denitrification(hostsim):  # tests/test_land.py:1012-1021 (temperature factor, then denitrification land.py:2096-2152)
    r = run_process(hostsim, ["temperature_factor", "denitrification"])
    assert only(r.vq("mb_in")) and only(r.vq("mb_out"), nitrate=0.03295446191742672)
    assert r.vq("storage")["nitrate"] == pytest.approx(2.5 - 0.03295446191742672, abs=1e-15)
    assert r.pool("dis_inorg")["N"] == pytest.approx(4.1 - 0.03295446191742672, abs=1e-15)


def test_desorption(hostsim):  # tests/test_land.py:1023-1039 (adsorption land.py:2154-2265, Freundlich / Newton-Raphson)
    def prepare(s):
        s.nutrient_pool.adsorbed_inorganic_pool.storage["P"] = 1e6

    r = run_process(hostsim, ["adsorption"], prepare=prepare)
    assert only(r.vq("mb_out")) and only(r.vq("mb_in"), phosphate=29534.602916697728)
    assert r.pool("ads_inorg") == {"N": pytest.approx(0.9, abs=1e-15), "P": pytest.approx(970465.3970833023, rel=1e-15)}
    assert r.pool("dis_inorg") == {"N": pytest.approx(4.1, abs=1e-15), "P": pytest.approx(29545.602916697728, rel=1e-15)}
    assert r.vq("storage")["phosphate"] == pytest.approx(11 + 29534.602916697728, rel=1e-15)


def test_adsorption(hostsim):  # tests/test_land.py:1041-1069 (fertiliser land.py:1649-1679, then adsorption)
    d = pd.Timestamp("2000-05-01")
    data = {("srp-fertiliser", d.to_period("M")): 1000.0}
    r = run_process(hostsim, ["fertiliser", "adsorption"], surface_data=data)
    assert only(r.vq("mb_out"), phosphate=1.5761692429741743)
    assert r.pool("ads_inorg") == {"N": pytest.approx(0.9, abs=1e-15), "P": pytest.approx(2.7761692429741744, rel=1e-14)}
    assert r.pool("dis_inorg") == {"N": pytest.approx(4.1, abs=1e-15), "P": pytest.approx(1509.4238307570258, rel=1e-14)}


def test_crop_cover(hostsim):  # tests/test_land.py:627-642 (calc_crop_cover land.py:1514-1572)
    def prepare(s):
        s.days_after_sow = pd.Timestamp("2000-05-01").dayofyear - s.sowing_day

    from wsi_b200.backend import layout as L

    r = run_process(hostsim, ["crop_cover"], prepare=prepare)
    plane = lambda f: r.g["surf"][L.SURF.offset(f, r.fm.V), 0]  # noqa: E731
    assert plane("crop_cover") == 0.17647058823529413
    assert plane("ground_cover") == 0.058823529411764705
    assert plane("days_after_sow") == 122 - 91 + 1


def test_fertiliser(hostsim):  # tests/test_land.py:717-744 (fertiliser land.py:1649-1679, allocate_fertiliser nutrient_pool.py:314-328)
    m = pd.Timestamp("2000-05-01").to_period("M")
    data = {("nhx-fertiliser", m): 5.0, ("noy-fertiliser", m): 2.0, ("srp-fertiliser", m): 3.0}
    r = run_process(hostsim, ["fertiliser"], surface_data=data)
    assert only(r.vq("mb_out")) and only(r.vq("mb_in"), phosphate=3 * 1.5, nitrate=2 * 1.5, ammonia=5 * 1.5)
    assert r.pool("dis_inorg") == {"N": pytest.approx(4.1 + (2 + 5) * 1.5, rel=1e-15), "P": 11 + 3 * 1.5}


def test_manure(hostsim):  # tests/test_land.py:746-790 (manure land.py:1681-1710, allocate_manure nutrient_pool.py:267-290)
    m = pd.Timestamp("2000-05-01").to_period("M")
    data = {("nhx-manure", m): 5.0, ("noy-manure", m): 2.0, ("srp-manure", m): 3.0}
    frac = {}

    def prepare(s):
        frac["di"] = dict(s.nutrient_pool.fraction_manure_to_dissolved_inorganic)
        frac["fast"] = dict(s.nutrient_pool.fraction_manure_to_fast)

    r = run_process(hostsim, ["manure"], surface_data=data, prepare=prepare)
    assert only(r.vq("mb_out"))
    assert only(r.vq("mb_in"), phosphate=3 * 1.5 * frac["di"]["P"], nitrate=2 * 1.5 * frac["di"]["N"], ammonia=5 * 1.5 * frac["di"]["N"])
    assert r.pool("dis_inorg") == {"N": pytest.approx(4.1 + (2 + 5) * 1.5 * frac["di"]["N"], rel=1e-15),
                                   "P": pytest.approx(11 + 3 * 1.5 * frac["di"]["P"], rel=1e-15)}
    assert r.pool("fast") == {"N": pytest.approx(2 + (2 + 5) * 1.5 * frac["fast"]["N"], rel=1e-15),
                              "P": pytest.approx(4 + 3 * 1.5 * frac["fast"]["P"], rel=1e-15)}


def test_soil_pool(hostsim):  # tests/test_land.py:839-910 (soil_pool_transformation land.py:1744-1811, nutrient_pool.py:396-450)
    par = {}

    def prepare(s):
        for k in ("minfpar", "disfpar", "immobdpar"):
            par[k] = dict(getattr(s.nutrient_pool, k))
        par["satact"] = s.satact

    r = run_process(hostsim, ["temperature_factor", "moisture_factor", "soil_pool"], prepare=prepare)
    # storage temperature 11 -> 2 ** ((11 - 20) / 10) (land.py:1813-1831); 0.32 / 1.5 m of water is above the field
    # capacity of 0.2 * 0.5 m -> satact (land.py:1833-1856)
    cf = 2 ** ((11 - 20) / 10) * par["satact"]
    fast = {"N": ISOIL["org-nitrogen"], "P": ISOIL["org-phosphorus"]}
    dis = {"N": IVOL["nitrate"] + IVOL["nitrite"] + IVOL["ammonia"], "P": IVOL["phosphate"]}
    miner = {k: par["minfpar"][k] * fast[k] * cf for k in "NP"}
    fast = {k: fast[k] - miner[k] for k in "NP"}
    dis = {k: dis[k] + miner[k] for k in "NP"}
    disso = {k: par["disfpar"][k] * fast[k] * cf for k in "NP"}
    immob = {k: par["immobdpar"][k] * dis[k] * cf for k in "NP"}
    d1, d2 = {"org_phosphorus": disso["P"], "org_nitrogen": disso["N"]}, {}
    (d1 if miner["P"] - immob["P"] > 0 else d2)["phosphate"] = abs(miner["P"] - immob["P"])
    nh, no = IVOL["ammonia"], IVOL["nitrate"]
    tgt = d1 if miner["N"] - immob["N"] > 0 else d2
    tgt["nitrate"] = abs(miner["N"] - immob["N"]) * no / (no + nh)
    tgt["ammonia"] = abs(miner["N"] - immob["N"]) * nh / (no + nh)
    assert only(r.vq("mb_in"), **d1) and only(r.vq("mb_out"), **d2)
