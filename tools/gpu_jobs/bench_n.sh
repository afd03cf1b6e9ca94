#!/bin/bash
# gpurun --gpus N job: the bench line at N ranks exactly as the driver launches it.  usage: bench_n.sh <N> [bench args]
N=${1:-8}; shift
mkdir -p gpurun_out
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 \
    bench.py --gpus $N "$@" > gpurun_out/bn_$N.json 2> gpurun_out/bn_$N.err ) 2>&1 | grep real
python - <<PY
import json
d=json.load(open("gpurun_out/bn_$N.json"))
print("n_gpus", d["n_gpus"], "ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"] and d["e2e"]["value"], "launches", d["gpu_launches"])
for k, v in (d.get("secondary") or {}).items():
    print(k, v.get("ms_per_step"), v.get("value"), v.get("n_gpus_used"))
PY
tail -3 gpurun_out/bn_$N.err
