"""Per-phase clock64() totals of the tail kernel (needs a build with -DWSI_NET_DEBUG)."""
import sys, ctypes as C
sys.path.insert(0, "/root/repo")
import torch
from wsi_b200 import rivernet
from wsi_b200.backend.engine import GpuRunner
from wsi_b200.backend.probes import ProbeSet
N = int(sys.argv[1])
fm = rivernet.build_dendritic(N, T=8, names=False)
ps = ProbeSet(fm).add_flows([0])
run = GpuRunner(fm, probes=ps.probes, rec_t0=0, rec_T=8, engine="network")
run.run(0, 8); run.synchronize()
out = (C.c_ulonglong * 16)()
print(run.lib.wsi_net_debug(out), [int(x) // 8 for x in out[:11]], "cycles per step; levels", run.buffers.net["n_levels"])
