#!/bin/bash
# gpurun job: host topology of the GPU box (NUMA nodes, GPU affinity) and pinned D2H bandwidth per NUMA node
mkdir -p gpurun_out
{
lscpu | grep -i "numa\|socket\|model name\|^CPU(s)"
nvidia-smi topo -m
for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q "^0x0302\|^0x0300" $d/class 2>/dev/null; then echo "$d numa=$(cat $d/numa_node) vendor=$(cat $d/vendor)"; fi; done
python - <<'PY'
import os, time, torch
print("affinity", len(os.sched_getaffinity(0)), "cpus")
p = torch.cuda.get_device_properties(0)
print("pci", getattr(p, "pci_bus_id", None), getattr(p, "pci_domain_id", None), getattr(p, "pci_device_id", None))
nodes = sorted(int(x[4:]) for x in os.listdir("/sys/devices/system/node") if x.startswith("node"))
def cpus(n):
    out = set()
    for part in open(f"/sys/devices/system/node/node{n}/cpulist").read().strip().split(","):
        a, _, b = part.partition("-"); out |= set(range(int(a), int(b or a) + 1))
    return out
allowed = os.sched_getaffinity(0)
dev = torch.zeros(10_400_000 // 8, dtype=torch.float64, device="cuda")
for n in nodes:
    c = cpus(n) & allowed
    if not c: print("node", n, "no allowed cpus"); continue
    os.sched_setaffinity(0, c)
    host = torch.empty(dev.shape, dtype=torch.float64).pin_memory()
    host.zero_()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200): host.copy_(dev, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("node", n, "cpus", len(c), "D2H GB/s", 200 * 10.4e-3 / dt)
    os.sched_setaffinity(0, allowed)
PY
} > gpurun_out/topo.log 2>&1
cat gpurun_out/topo.log
