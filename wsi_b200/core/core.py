"""Host-side VQIP helpers with the reference's WSIObj method names
(wsimod/core/core.py:13-429).  They exist so user code that builds or inspects
VQIP dicts keeps working; the simulation itself never calls them -- the hot path
runs in the CUDA kernels (wsi_b200/csrc/wsi_vq.cuh)."""
from __future__ import annotations

from . import constants


class WSIObj:
    def __init__(self):
        self.empty_vqip_predefined = dict.fromkeys(list(constants.POLLUTANTS) + ["volume"], 0)

    # -- construction ------------------------------------------------------------------
    def empty_vqip(self):
        return dict(self.empty_vqip_predefined)

    def copy_vqip(self, t):
        return dict(t)

    # -- totals algebra (core.py:83-116, 150-178, 209-223, 251-284, 305-328) --------------
    def sum_vqip(self, t1, t2):
        out = dict(t1)
        out["volume"] = t1["volume"] + t2["volume"]
        for p in constants.ADDITIVE_POLLUTANTS:
            out[p] = t1[p] + t2[p]
        if out["volume"] > 0:
            for p in constants.NON_ADDITIVE_POLLUTANTS:
                out[p] = (t2[p] * t2["volume"] + t1[p] * t1["volume"]) / out["volume"]
        return out

    def extract_vqip(self, t1, t2):
        out = dict(t1)
        for p in list(constants.ADDITIVE_POLLUTANTS) + ["volume"]:
            out[p] = t1[p] - t2[p]
        return out

    def v_distill_vqip(self, t, v):
        out = dict(t)
        out["volume"] = t["volume"] - v
        return out

    def v_change_vqip(self, t, v):
        out = dict(t)
        if t["volume"] > 0:
            ratio = v / t["volume"]
            out["volume"] = t["volume"] * ratio
            for p in constants.ADDITIVE_POLLUTANTS:
                out[p] = t[p] * ratio
        else:
            out["volume"] = v
        return out

    def ds_vqip(self, t, t_):
        out = self.empty_vqip()
        for p in list(constants.ADDITIVE_POLLUTANTS) + ["volume"]:
            out[p] = t[p] - t_[p]
        return out

    def compare_vqip(self, t1, t2):
        return all(abs(t1[k] - t2[k]) <= constants.FLOAT_ACCURACY for k in t1)

    # -- concentration-space variants (core.py:56-81, 118-148, 180-207, 225-249, 286-352) ---
    def blend_vqip(self, c1, c2):
        out = self.empty_vqip()
        out["volume"] = c1["volume"] + c2["volume"]
        if out["volume"] > 0:
            for p in constants.POLLUTANTS:
                out[p] = (c1[p] * c1["volume"] + c2[p] * c2["volume"]) / out["volume"]
        return out

    def concentration_to_total(self, c):
        out = dict(c)
        for p in constants.ADDITIVE_POLLUTANTS:
            out[p] = c[p] * c["volume"]
        return out

    def total_to_concentration(self, t):
        out = dict(t)
        for p in constants.ADDITIVE_POLLUTANTS:
            out[p] = t[p] / t["volume"]
        return out

    def extract_vqip_c(self, c1, c2):
        out = dict(c1)
        t1, t2 = self.concentration_to_total(c1), self.concentration_to_total(c2)
        out["volume"] = t1["volume"] - t2["volume"]
        if out["volume"] > 0:
            for p in constants.ADDITIVE_POLLUTANTS:
                out[p] = (t1[p] - t2[p]) / out["volume"]
        return out

    def v_distill_vqip_c(self, c, v):
        d = self.empty_vqip()
        d["volume"] = -v
        out = self.blend_vqip(dict(c), d)
        for p in constants.NON_ADDITIVE_POLLUTANTS:
            out[p] = c[p]
        return out

    def v_change_vqip_c(self, c, v):
        out = dict(c)
        out["volume"] = v
        return out

    def ds_vqip_c(self, c, c_):
        out = self.empty_vqip()
        out["volume"] = c["volume"] - c_["volume"]
        for p in constants.ADDITIVE_POLLUTANTS:
            out[p] = c["volume"] * c[p] - c_["volume"] * c_[p]
        return out


class DecayObj(WSIObj):
    """Holds decay parameters {pollutant: {constant, exponent}} (core.py:432-472); the decay
    itself (core.py:494-521) runs on the device."""

    def __init__(self, decays):
        self.decays = decays
        WSIObj.__init__(self)
        self.total_decayed = self.empty_vqip()
