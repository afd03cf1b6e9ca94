"""Small network-engine run used as the target of ncu captures (python tools/net_ncu_target.py N)."""
import sys, os
sys.path.insert(0, "/root/repo")
import torch
from wsi_b200 import rivernet
from wsi_b200.backend.engine import GpuRunner
from wsi_b200.backend.probes import ProbeSet
N = int(sys.argv[1])
fm = rivernet.build_dendritic(N, T=4, names=False)
ps = ProbeSet(fm).add_flows([0])
run = GpuRunner(fm, probes=ps.probes, rec_t0=0, rec_T=4, engine="network")
run.run(0, 4); run.synchronize()
print("done")
