#!/bin/bash
# gpurun job: GPU parity tests, the reference arm and the full default bench line (all secondary blocks).
tag=${1:-run}; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${tag}_ref.json 2> gpurun_out/${tag}_ref.err; cut -c 1-400 gpurun_out/${tag}_ref.json
( time python bench.py --steps 20 --warmup 5 "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err ) 2>&1 | grep real
python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench.json"))
print("ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"] and d["e2e"]["value"], "cpu", d["cpu_baseline"] and (d["cpu_baseline"]["kind"], d["cpu_baseline"]["value"]))
print(d["roofline"]["kernel_ms"])
print(json.dumps(d.get("secondary"), indent=1)[:3000])
PY
tail -3 gpurun_out/${tag}_bench.err
