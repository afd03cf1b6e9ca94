"""The CPU oracle (oracle/wsi_oracle.c) against golden vectors produced by RUNNING THE
REFERENCE (tests/golden/make_golden.py) and against known answers of the reference's own
unit tests.  This is what pins the oracle."""
import numpy as np
import pytest

from parity_util import compare_blocks, compare_mb, load_golden

GOLDENS = ["quickstart", "quickstart_simple", "quickstart_p5", "quickstart_p14", "fixture3607", "config2", "rivernet", "rivertree", "config3", "hourly"]


@pytest.mark.parametrize("name", GOLDENS)
def test_oracle_matches_reference_outputs(name):
    from oracle import oracle
    from wsi_b200.backend.probes import ProbeSet

    fm, ref = load_golden(name)
    ps = ProbeSet(fm).add_standard()
    run = oracle.run(fm, probes=ps.probes, record_mb=True, n_threads=2)
    blocks = ps.split(run.out_probe)
    ok, reports = compare_blocks(blocks, ref, fm, rtol=1e-12)
    ok2, rep = compare_mb(run.out_mb, ref["mb"], fm, rtol=1e-12)
    assert ok and ok2, "\n".join(reports + [rep])
    # the oracle follows the reference operation for operation: flows are bit-identical
    flows_ref = ref["flows"] if fm.R > 1 else ref["flows"][..., None]
    assert np.array_equal(blocks["flows"], flows_ref.reshape(blocks["flows"].shape))


@pytest.mark.parametrize("name", ["config2", "fixture3607"])
def test_oracle_mass_balance_advisories_match_reference(name):
    """The reference prints one advisory line per (object, component) whose balance does not
    close (core.py:419-427); the oracle's violation count must equal the number of lines."""
    from oracle import oracle

    fm, ref = load_golden(name)
    run = oracle.run(fm)
    if name == "config2":
        assert int(run.mb_count.sum()) == int(ref["n_advisories"][0])
    else:  # fixture prints river/evaporation advisories too, which are not mass-balance errors
        assert int(run.mb_count.sum()) <= int(ref["n_advisories"][0])


def _kept_flows(fm, ref, out_probe, ps):
    idx = [fm.arc_names.index(str(a)) for a in ref["kept_arcs"]]
    return ps.split(out_probe)["flows"][:, idx]


def test_config3_ensemble_members_are_the_reference_built_members():
    """wsi_b200.ensemble.build_config3 perturbs the flat planes of the Oxford member; its first 16
    members must be the 16 members tests/golden/make_golden.py built object by object WITH THE
    REFERENCE (same factors), and the oracle must reproduce the reference's flows on all of them."""
    from oracle import oracle
    from wsi_b200 import ensemble as E
    from wsi_b200.backend.probes import ProbeSet

    template, _ = load_golden("config3")
    g16, ref = load_golden("config3_members16")
    fm = E.build_config3(16, template)
    for g in fm.groups:
        if fm.groups[g].size:
            np.testing.assert_allclose(fm.groups[g], g16.groups[g], rtol=1e-14, atol=0, err_msg=g)
    f = E.config3_factors(16)
    assert np.all(f[0] == 1.0) and len(np.unique(f[1:, 0])) == 15 and f[1:].min() >= 0.8 and f[1:].max() <= 1.25
    ps = ProbeSet(g16).add_flows()
    run = oracle.run(g16, probes=ps.probes, n_threads=4)
    got = _kept_flows(g16, ref, run.out_probe, ps)
    assert np.array_equal(got, ref["flows_kept"].reshape(got.shape))


def test_oracle_chunked_run_equals_single_run():
    """State carried in the SoA tensors is complete: [0,T) == [0,a) + [a,T)."""
    from oracle import oracle
    from wsi_b200.backend.probes import ProbeSet

    fm, _ = load_golden("rivernet")
    ps = ProbeSet(fm).add_flows()
    full = oracle.run(fm, probes=ps.probes, t0=0, t1=120)
    a = oracle.run(fm, probes=ps.probes, t0=0, t1=47)
    fm2 = fm.copy()
    fm2.groups = {k: v.copy() for k, v in a.groups.items()}
    b = oracle.run(fm2, probes=ps.probes, t0=47, t1=120)
    assert np.array_equal(full.out_probe[:47], a.out_probe)
    assert np.array_equal(full.out_probe[47:], b.out_probe)
    for g in full.groups:
        assert np.array_equal(full.groups[g], b.groups[g]), g


def test_reference_unit_test_known_answers():
    """tests/test_model.py:189-295 (reference): one day of the quickstart model gives
    flows[-1]['flow'] == 0.05 and 0.03 left on the urban surface."""
    from oracle import oracle
    from wsi_b200.backend.probes import ProbeSet
    from wsi_b200.core import constants
    from wsi_b200.orchestration.model import Model, to_datetime

    constants.set_default_pollutants()
    constants.FLOAT_ACCURACY = 1e-11
    date = to_datetime("2000-01-01")
    land_inputs = {("temperature", date): 10, ("precipitation", date): 0.01, ("et0", date): 0.002}
    m = Model()
    m.dates = [date]
    m.add_nodes([
        {"type_": "Sewer", "capacity": 0.05, "name": "my_sewer"},
        {"type_": "Land", "data_input_dict": land_inputs, "name": "my_land", "surfaces": [
            {"type_": "ImperviousSurface", "surface": "urban", "area": 10, "pollutant_load": {"phosphate": 1e-7}},
            {"type_": "PerviousSurface", "surface": "rural", "area": 100, "depth": 0.5,
             "pollutant_load": {"phosphate": 1e-7}}]},
        {"type_": "Groundwater", "area": 100, "capacity": 100, "name": "my_groundwater"},
        {"type_": "Node", "name": "my_river"},
        {"type_": "Waste", "name": "my_outlet"},
    ])
    arcs = [("urban_drainage", "my_land", "my_sewer"), ("percolation", "my_land", "my_groundwater"),
            ("runoff", "my_land", "my_river"), ("storm_outflow", "my_sewer", "my_river"),
            ("baseflow", "my_groundwater", "my_river"), ("catchment_outflow", "my_river", "my_outlet")]
    m.add_arcs([{"type_": "Arc", "name": a, "in_port": i, "out_port": o} for a, i, o in arcs])
    fm = m.flatten()
    ps = ProbeSet(fm).add_standard()
    run = oracle.run(fm, probes=ps.probes)
    out = ps.split(run.out_probe)
    assert out["flows"][-1, -1, 0, 0] == 0.05  # catchment_outflow volume
    assert out["surf_storage"][-1, 0, 0, 0] == 0.03  # urban surface storage


def test_vqip_algebra_known_answers():
    """tests/test_core.py:32-44 (reference): sum_vqip blends non-additive pollutants to 40/3;
    the host-side mirror of WSIObj reproduces the hand numbers."""
    from wsi_b200.core import constants
    from wsi_b200.core.core import WSIObj

    constants.set_simple_pollutants()
    try:
        obj = WSIObj()
        a = {"volume": 10, "phosphate": 0.0001, "temperature": 15}
        b = {"volume": 5, "phosphate": 0.00005, "temperature": 10}
        s = obj.sum_vqip(a, b)
        assert s["volume"] == 15 and s["temperature"] == 40 / 3
        assert abs(s["phosphate"] - 0.00015) < 1e-19
        assert obj.extract_vqip(a, b)["volume"] == 5
        half = obj.v_change_vqip(a, 5)
        assert half["volume"] == 5 and half["phosphate"] == 0.00005 and half["temperature"] == 15
        assert obj.ds_vqip(a, b) == {"volume": 5, "phosphate": 0.00005, "temperature": 0}
    finally:
        constants.set_default_pollutants()
