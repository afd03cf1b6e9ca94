#!/bin/bash
# gpurun job: what the round's evidence files are made from -- GPU parity tests, smoke(), the reference arm, the
# default bench line (as the driver runs it), then the launch list of a few timesteps under ncu.
tag=${1:-final}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -2 gpurun_out/${tag}_smoke.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${tag}_ref.json 2> gpurun_out/${tag}_ref.err; cut -c 1-300 gpurun_out/${tag}_ref.json; echo
( time python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err ) 2>&1 | grep real
python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench.json"))
print("ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"] and d["e2e"]["value"], "cpu", d["cpu_baseline"] and (d["cpu_baseline"]["kind"], d["cpu_baseline"]["value"]))
print(d["roofline"]["kernel_ms"], d["gpu_launches"], d["clocks"])
for k, v in (d.get("secondary") or {}).items():
    print(k, v.get("ms_per_step"), v.get("value"), v.get("e2e", {}).get("value") if isinstance(v.get("e2e"), dict) else None)
PY
tail -3 gpurun_out/${tag}_bench.err
bash tools/gpu_jobs/launch_list.sh ${tag}_launches | tail -12
