#!/bin/bash
# gpurun job: GPU parity tests, then a short bench run.  usage: bash tools/gpu_jobs/tests_and_bench.sh <tag> [bench args]
tag=${1:-run}; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log
python bench.py --steps 100 --warmup 5 --no-cpu "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 1800 gpurun_out/${tag}_bench.json
