"""Parity metric and golden-fixture helpers shared by the CPU and GPU tests.

Tolerance (BASELINE.json north_star: 1e-9 relative):

    |a - b| <= RTOL * max(|a|, |b|) + atol,
    atol = FLOAT_ACCURACY * 10**int(log10(scale))

``scale`` is the largest magnitude the compared quantity (one column of a named
block, e.g. one pollutant on one arc) reaches in the reference series -- the
reference's own notion of "zero at this magnitude" (wsimod/core/core.py:408-419).
Exact zeros in one run are O(1e-16 .. 1e-11) cancellation residue in another
(SURVEY 4.4), so a pure relative test is meaningless there; the pure-relative
worst case over entries above the floor is reported alongside.
"""
from __future__ import annotations

import os

import numpy as np

RTOL = 1e-9
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    from wsi_b200.backend.flatten import FlatModel

    path = os.path.join(GOLDEN_DIR, name + ".npz")
    d = dict(np.load(path, allow_pickle=False))
    fm = FlatModel.from_npz_dict(d)
    ref = {k[4:]: v for k, v in d.items() if k.startswith("ref_")}
    return fm, ref


def ref_with_R(ref_arr, R):
    """Golden arrays carry a trailing replica axis only when R > 1."""
    a = np.asarray(ref_arr)
    return a if R > 1 else a[..., None]


def compare(name, got, ref, float_accuracy, rtol=RTOL, time_axis=0):
    """Return (ok, report) for two arrays of identical shape ``[T, ...]``."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, f"{name}: shape {got.shape} != {ref.shape}"
    if got.size == 0:
        return True, f"{name}: empty"
    scale = np.max(np.abs(ref), axis=time_axis, keepdims=True)
    with np.errstate(divide="ignore"):
        mag = np.where(scale > float_accuracy, 10.0 ** np.trunc(np.log10(np.maximum(scale, 1e-300))), 1.0)
    atol = float_accuracy * mag
    err = np.abs(got - ref)
    big = np.maximum(np.abs(got), np.abs(ref))
    bad = err > rtol * big + atol
    nan_bad = np.isnan(got) != np.isnan(ref)
    bad |= nan_bad
    above = big > atol * 1e3
    rel = np.where(above, err / np.maximum(big, 1e-300), 0.0)
    worst_rel = float(rel.max()) if rel.size else 0.0
    report = f"{name}: worst pure-relative {worst_rel:.3e} (entries above floor), max abs err {float(np.nanmax(err)):.3e}"
    if bad.any():
        idx = np.unravel_index(np.argmax(np.where(bad, err / (rtol * big + atol), 0)), bad.shape)
        report += (
            f"; {int(bad.sum())} of {bad.size} entries out of tolerance, worst at {idx}: "
            f"got {got[idx]!r} ref {ref[idx]!r}"
        )
    return (not bad.any()), report


def compare_blocks(got_blocks, ref, fm, names=None, rtol=RTOL):
    """Compare every named block of ``ProbeSet.split`` output with the golden dict."""
    ok_all, reports = True, []
    for name, got in got_blocks.items():
        if names is not None and name not in names:
            continue
        r = ref_with_R(ref[name], fm.R)
        r = r.reshape(got.shape)
        ok, rep = compare(name, got, r, fm.float_accuracy, rtol)
        ok_all &= ok
        reports.append(rep)
    return ok_all, reports


def compare_mb(out_mb, ref_mb, fm, rtol=RTOL):
    """Mass-balance terms: only volume + additive planes are defined (core.py:400-404)."""
    got = np.asarray(out_mb)  # [T, n_nodes, 3, V, R]
    r = ref_with_R(ref_mb, fm.R)
    planes = [0] + [fm.pollutants.index(p) + 1 for p in fm.additive]
    return compare("mass_balance", got[:, :, :, planes, :], r[:, :, :, planes, :], fm.float_accuracy, rtol)
