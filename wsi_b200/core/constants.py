"""Module-level constants with the reference's names and values
(wsimod/core/constants.py).  As in the reference the pollutant lists and
FLOAT_ACCURACY are *mutable module globals*: Model.load rewrites them from the YAML
file (model.py:232-241) and the flattener reads them when a run starts."""

M3_S_TO_ML_D = 86.4
MM_KM2_TO_ML = 1e-3 * 1e6 * 1e3 * 1e-6
MM_M2_TO_ML = 1e-3 * 1e3 * 1e-6
MM_M2_TO_SIM_VOLUME = MM_M2_TO_ML
MM_M2_TO_M3 = 1e-3
ML_TO_M3 = 1000
PCT_TO_PROP = 1 / 100
L_TO_ML = 1e-6
FLOAT_ACCURACY = 1e-11
UNBOUNDED_CAPACITY = 1e15
MAX_INFLOW = 1e10
DT_DAYS = 1
_DEFAULT_POLLUTANTS = (
    "do", "org-phosphorus", "phosphate", "ammonia", "solids", "bod", "cod", "ph", "temperature",
    "nitrate", "nitrite", "org-nitrogen",
)
_DEFAULT_NON_ADDITIVE = ("do", "temperature", "ph")
POLLUTANTS = list(_DEFAULT_POLLUTANTS)
NON_ADDITIVE_POLLUTANTS = list(_DEFAULT_NON_ADDITIVE)
ADDITIVE_POLLUTANTS = [p for p in _DEFAULT_POLLUTANTS if p not in _DEFAULT_NON_ADDITIVE]
NUTRIENTS = ["N", "P"]
PCT_GARDENS = 0.1
PI = 3.141592653589793
PER_DAY_TO_PER_SECOND = 1 / (60 * 60 * 24)
PER_DAY_TO_PER_HOUR = 1 / 24
M3_S_TO_ML_H = M3_S_TO_ML_D * PER_DAY_TO_PER_HOUR
MAXITER = 5
M3_S_TO_M3_DT = 86400
M_S_TO_M_DT = 86400
UG_L_TO_KG_M3 = 1e-6
MG_L_TO_KG_M3 = 1e-3
KG_M3_TO_MG_L = 1e3
G_TO_KG = 1e-3
KG_TO_MG = 1e6
DECAY_REFERENCE_TEMPERATURE = 20
PER_30MONTH_TO_PER_DAY = 1 / 30
DAYS_IN_YEAR = 365
L_S_TO_M3_D = 1e-3 / PER_DAY_TO_PER_SECOND
M_TO_MM = 1e3
MM_TO_M = 1e-3
KM_TO_M = 1e3
M_TO_KM = 1e-3
MJ_M2_TO_CAL_CM2 = 23.889
KPA_TO_MBAR = 10
HPA_TO_KPA = 0.1
KM2_TO_HA = 100
KM2_TO_M2 = 1e6
HA_TO_M2 = 1e4
M2_TO_KM2 = 1e-6
G_M2_TO_KG_KM2 = 1e3
G_M2_TO_KG_M2 = 1e-3
MGMM_L_TO_KG_KM2 = 1e-6 * 1e3 * 1e-3 * 1e6
KG_M2_TO_KG_KM2 = 1e6
KG_M3_TO_KG_KM3 = 1e9
D_TO_S = 3600 * 24


def set_simple_pollutants():
    """Two-pollutant configuration (constants.py:85-89)."""
    global POLLUTANTS, ADDITIVE_POLLUTANTS, NON_ADDITIVE_POLLUTANTS
    POLLUTANTS = ["phosphate", "temperature"]
    ADDITIVE_POLLUTANTS = ["phosphate"]
    NON_ADDITIVE_POLLUTANTS = ["temperature"]


def set_default_pollutants():
    """Twelve default pollutants (constants.py:92-132)."""
    global POLLUTANTS, ADDITIVE_POLLUTANTS, NON_ADDITIVE_POLLUTANTS
    POLLUTANTS = list(_DEFAULT_POLLUTANTS)
    NON_ADDITIVE_POLLUTANTS = list(_DEFAULT_NON_ADDITIVE)
    ADDITIVE_POLLUTANTS = [p for p in _DEFAULT_POLLUTANTS if p not in _DEFAULT_NON_ADDITIVE]
