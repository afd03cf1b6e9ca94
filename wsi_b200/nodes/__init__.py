from .catchment import Catchment  # noqa: F401
from .demand import Demand, ResidentialDemand  # noqa: F401
from .distribution import Distribution, UnlimitedDistribution  # noqa: F401
from .land import Land  # noqa: F401
from .nodes import NODES_REGISTRY, Node  # noqa: F401
from .sewer import Sewer  # noqa: F401
from .storage import Groundwater, QueueGroundwater, Reservoir, River, RiverReservoir, Storage  # noqa: F401
from .waste import Waste  # noqa: F401
from .wtw import FWTW, WTW, WWTW  # noqa: F401
