#!/bin/bash
# gpurun job: short bench run (no CPU leg).  usage: bash tools/gpu_jobs/bench_only.sh <tag> [bench args]
tag=${1:-run}; shift
mkdir -p gpurun_out
python bench.py --steps 100 --warmup 5 --no-cpu "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench.json"))
print("ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"] and d["e2e"]["value"])
print(d["roofline"]["kernel_ms"])
PY
tail -3 gpurun_out/${tag}_bench.err
