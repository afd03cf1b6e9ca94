#!/bin/bash
# gpurun --gpus 2 job: the 2-GPU NCCL test, then bench at N=2 with and without the in-region all-reduce
mkdir -p gpurun_out
python -m pytest tests/test_distributed_gpu.py -x -q 2>&1 | tail -2
for nr in 0 1; do
WSI_BENCH_NO_REDUCE=$nr python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2955$nr bench.py --gpus 2 --steps 40 --warmup 8 --secondary "" --no-e2e > gpurun_out/b2_$nr.json 2> gpurun_out/b2_$nr.err
python -c "
import json; d=json.load(open('gpurun_out/b2_$nr.json')); print('no_reduce=$nr', 'ms/step', d['ms_per_step'], 'value', d['value'])"
done
