#!/bin/bash
# gpurun job: short bench, then ncu --set full of one kernel.  usage: bash tools/gpu_jobs/bench_and_ncu.sh <tag> <regex> <skip> <count>
tag=${1:-run}; rx=${2:-wsi_step_kernel}; skip=${3:-4}; cnt=${4:-1}
bash tools/gpu_jobs/bench_only.sh $tag --no-e2e
bash tools/gpu_jobs/ncu_step.sh ${tag}_prof "$rx" $skip $cnt | tail -2
