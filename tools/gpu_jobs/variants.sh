#!/bin/bash
# gpurun job: bench every variant library under wsi_b200/csrc/variants (tools/build_variants.py).
# usage: bash tools/gpu_jobs/variants.sh <tag> [bench args]
tag=${1:-var}; shift
mkdir -p gpurun_out
for so in wsi_b200/csrc/variants/libwsi_*.so; do
  name=$(basename $so .so)
  WSI_B200_LIB=$so python bench.py --steps 60 --warmup 5 --no-cpu --no-e2e "$@" > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${tag}_${name}.json"))
    print("${name}", "ms/step %.4f" % d["ms_per_step"], "frac %.4f" % d["roofline"]["frac"], [round(v, 4) for v in d["roofline"]["kernel_ms"].values()])
except Exception as e:
    print("${name}", "FAILED", e)
PY
done
