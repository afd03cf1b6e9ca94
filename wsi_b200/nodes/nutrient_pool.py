"""NutrientPool parameters and the five N/P stores of a GrowingSurface
(reference: wsimod/nodes/nutrient_pool.py:11-136).  The transformations run in the
fused surface kernel (wsi_b200/csrc/wsi_land.cuh)."""
from __future__ import annotations


class NutrientStore:
    def __init__(self):
        self.storage = {"N": 0, "P": 0}


class NutrientPool:
    def __init__(
        self,
        fraction_dry_n_to_dissolved_inorganic=0.9,
        degrhpar=None,
        dishpar=None,
        minfpar=None,
        disfpar=None,
        immobdpar=None,
        fraction_manure_to_dissolved_inorganic=None,
        fraction_residue_to_fast=None,
    ):
        # fresh dicts per instance (the reference shares mutable defaults between pools,
        # SURVEY 7.3 #16; sharing is not reproduced)
        self.fraction_dry_n_to_dissolved_inorganic = fraction_dry_n_to_dissolved_inorganic
        self.degrhpar = dict(degrhpar or {"N": 7 * 1e-5, "P": 7 * 1e-6})
        self.dishpar = dict(dishpar or {"N": 7 * 1e-5, "P": 7 * 1e-6})
        self.minfpar = dict(minfpar or {"N": 0.00013, "P": 0.000003})
        self.disfpar = dict(disfpar or {"N": 0.000003, "P": 0.0000001})
        self.immobdpar = dict(immobdpar or {"N": 0.0056, "P": 0.2866})
        self.fraction_manure_to_dissolved_inorganic = dict(
            fraction_manure_to_dissolved_inorganic or {"N": 0.5, "P": 0.1}
        )
        self.fraction_residue_to_fast = dict(fraction_residue_to_fast or {"N": 0.1, "P": 0.1})
        self.temperature_dependence_factor = 0
        self.soil_moisture_dependence_factor = 0
        self.calculate_fraction_parameters()
        self.fast_pool = NutrientStore()
        self.humus_pool = NutrientStore()
        self.dissolved_inorganic_pool = NutrientStore()
        self.dissolved_organic_pool = NutrientStore()
        self.adsorbed_inorganic_pool = NutrientStore()
        self.pools = [
            self.fast_pool,
            self.humus_pool,
            self.dissolved_inorganic_pool,
            self.dissolved_organic_pool,
            self.adsorbed_inorganic_pool,
        ]

    def calculate_fraction_parameters(self):
        self.fraction_manure_to_fast = {k: 1 - v for k, v in self.fraction_manure_to_dissolved_inorganic.items()}
        self.fraction_residue_to_humus = {k: 1 - v for k, v in self.fraction_residue_to_fast.items()}
        self.fraction_dry_n_to_fast = 1 - self.fraction_dry_n_to_dissolved_inorganic

    def apply_overrides(self, overrides=None):
        overrides = overrides or {}
        self.fraction_dry_n_to_dissolved_inorganic = overrides.pop(
            "fraction_dry_n_to_dissolved_inorganic", self.fraction_dry_n_to_dissolved_inorganic
        )
        for key in ("fraction_residue_to_fast", "fraction_manure_to_dissolved_inorganic", "degrhpar", "dishpar",
                    "minfpar", "disfpar", "immobdpar"):
            getattr(self, key).update(overrides.pop(key, {}))
        self.calculate_fraction_parameters()
