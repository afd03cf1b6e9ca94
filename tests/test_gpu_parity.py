"""GPU parity: the CUDA path (through the C ABI) against the golden vectors produced by the
reference itself and against the CPU oracle on the same inputs."""
import numpy as np
import pytest

from parity_util import compare_blocks, compare_mb, load_golden

pytestmark = pytest.mark.gpu

GOLDENS = ["quickstart", "quickstart_simple", "fixture3607", "config2", "rivernet"]


def _run_gpu(fm, ps, record_mb=True, chunks=1):
    from wsi_b200.backend.engine import GpuRunner

    run = GpuRunner(fm, probes=ps.probes, rec_t0=0, rec_T=fm.T, record_mb=record_mb)
    edges = np.linspace(0, fm.T, chunks + 1).astype(int)
    for a, b in zip(edges[:-1], edges[1:]):
        run.run(int(a), int(b))
    run.synchronize()
    return run


@pytest.mark.parametrize("name", GOLDENS)
def test_golden_parity(name):
    from wsi_b200.backend.probes import ProbeSet

    fm, ref = load_golden(name)
    ps = ProbeSet(fm).add_standard()
    run = _run_gpu(fm, ps, chunks=3)
    blocks = ps.split(run.out_probe())
    ok, reports = compare_blocks(blocks, ref, fm)
    ok2, rep = compare_mb(run.out_mb(), ref["mb"], fm)
    print("\n".join(reports + [rep]))
    assert ok and ok2, "\n".join(reports + [rep])
    assert run.launch_count() > 0


@pytest.mark.parametrize("name", ["config2", "rivernet"])
def test_final_state_matches_oracle(name):
    """State left in the SoA tensors after the run equals the oracle's (scatter-back input)."""
    from oracle import oracle
    from wsi_b200.backend.probes import ProbeSet
    from parity_util import compare

    fm, _ = load_golden(name)
    ps = ProbeSet(fm).add_flows()
    run = _run_gpu(fm, ps, record_mb=False)
    ora = oracle.run(fm, probes=ps.probes)
    got = run.groups()
    for g, arr in ora.groups.items():
        if arr.size == 0:
            continue
        ok, rep = compare(g, got[g].T[None], arr.T[None], fm.float_accuracy, rtol=1e-9, time_axis=0)
        # time_axis=0 over a single row: scale per entry
        assert ok, rep
    # system totals and violation counts
    np.testing.assert_allclose(run.mb_sums(), ora.mb_sums, rtol=1e-8, atol=1e-6)
