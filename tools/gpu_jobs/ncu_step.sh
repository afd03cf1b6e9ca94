#!/bin/bash
# gpurun job: ncu --set full of selected kernels of one config-2 timestep.
# usage: bash tools/gpu_jobs/ncu_step.sh <tag> <kernel regex> <skip> <count>
tag=${1:-prof}; rx=${2:-wsi_step_kernel}; skip=${3:-4}; cnt=${4:-1}
mkdir -p gpurun_out
export WSI_B200_NO_GRAPH=1
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --set full --import-source on --clock-control none -k regex:"$rx" -s $skip -c $cnt -f -o gpurun_out/${tag} \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/${tag}_ncu.log 2>&1
tail -5 gpurun_out/${tag}_ncu.log
