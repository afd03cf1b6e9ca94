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

Non-additive lanes (temperature, ph, do: concentrations, not masses) are compared only
where the volume that carries them is itself above the noise floor (1e3 x the volume
tolerance): the "temperature of an empty tank" is whatever the last drop left behind, and
whether a store holds exactly 0 or 1e-17 m3 is decided by last-ulp rounding in the reference
itself (SURVEY 4.4 measured exactly this on the reference under a 1-ulp perturbation).
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


VQIP_BLOCKS = {"flows": 1, "land_tanks": 3, "surf_storage": 1, "tank_storage": 1, "sewer_storage": 1,
               "sewer_active": 1}


def nonadditive_mask(got, ref, fm, n_vq):
    """Boolean mask [T, n_obj, n_vq*V, R]: True where a non-additive lane rides on a volume
    below the noise floor in both runs (not comparable)."""
    V = fm.V
    T, n_obj, width, R = ref.shape
    assert width == n_vq * V
    g = got.reshape(T, n_obj, n_vq, V, R)
    r = ref.reshape(T, n_obj, n_vq, V, R)
    vol = np.maximum(np.abs(g[:, :, :, 0]), np.abs(r[:, :, :, 0]))  # [T, n_obj, n_vq, R]
    vscale = np.max(np.abs(r[:, :, :, 0]), axis=0, keepdims=True)
    with np.errstate(divide="ignore"):
        mag = np.where(vscale > fm.float_accuracy, 10.0 ** np.trunc(np.log10(np.maximum(vscale, 1e-300))), 1.0)
    floor = fm.float_accuracy * mag * 1e3
    dry = vol <= floor
    mask = np.zeros(g.shape, dtype=bool)
    for p in fm.non_additive:
        mask[:, :, :, fm.pollutants.index(p) + 1] = dry
    return mask.reshape(ref.shape)


def compare(name, got, ref, float_accuracy, rtol=RTOL, time_axis=0, ignore=None):
    """Return (ok, report) for two arrays of identical shape ``[T, ...]``."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, f"{name}: shape {got.shape} != {ref.shape}"
    if got.size == 0:
        return True, f"{name}: empty"
    if ignore is not None:
        got = np.where(ignore, ref, got)
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
        ignore = nonadditive_mask(got, r, fm, VQIP_BLOCKS[name]) if name in VQIP_BLOCKS else None
        ok, rep = compare(name, got, r, fm.float_accuracy, rtol, ignore=ignore)
        if ignore is not None and ignore.any():
            rep += f" [{int(ignore.sum())} non-additive entries on dry stores not compared]"
        ok_all &= ok
        reports.append(rep)
    return ok_all, reports


def compare_mb(out_mb, ref_mb, fm, rtol=RTOL):
    """Mass-balance terms: only volume + additive planes are defined (core.py:400-404)."""
    got = np.asarray(out_mb)  # [T, n_nodes, 3, V, R]
    r = ref_with_R(ref_mb, fm.R)
    planes = [0] + [fm.pollutants.index(p) + 1 for p in fm.additive]
    return compare("mass_balance", got[:, :, :, planes, :], r[:, :, :, planes, :], fm.float_accuracy, rtol)
