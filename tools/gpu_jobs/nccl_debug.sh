#!/bin/bash
# gpurun --gpus 2 job: the 2-GPU NCCL test's error text and the cost of the batched mass-balance all-reduce
mkdir -p gpurun_out
python -m pytest tests/test_distributed_gpu.py -x -q > gpurun_out/nccl_test.log 2>&1; grep -E "Error|error|assert|^E " gpurun_out/nccl_test.log | head -20
cat > /tmp/nccl_probe.py <<'PY'
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch, torch.distributed as dist
rank = int(os.environ['LOCAL_RANK']); torch.cuda.set_device(rank); dev = f'cuda:{rank}'
dist.init_process_group('nccl', device_id=torch.device(dev))
a = torch.zeros(8, 3, 13, dtype=torch.float64, device=dev); b = torch.zeros(8, dtype=torch.int32, device=dev)
for _ in range(3):
    dist.all_reduce(a); dist.all_reduce(b)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    dist.all_reduce(a); dist.all_reduce(b)
torch.cuda.synchronize()
if rank == 0: print('sync all_reduce pair: %.1f us' % ((time.perf_counter() - t0) / 20 * 1e6))
t0 = time.perf_counter()
ws = []
for _ in range(20):
    ws.append(dist.all_reduce(a, async_op=True)); ws.append(dist.all_reduce(b, async_op=True))
for w in ws: w.wait()
torch.cuda.synchronize()
if rank == 0: print('async all_reduce pair: %.1f us' % ((time.perf_counter() - t0) / 20 * 1e6))
dist.destroy_process_group()
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 /tmp/nccl_probe.py 2>&1 | grep -v "^\*\|OMP_NUM"
