#!/bin/bash
# gpurun job: ncu --set full of the kernels of one config-3 timestep (125 000 Oxford ensemble members).
tag=${1:-c3prof}
mkdir -p gpurun_out
export WSI_B200_NO_GRAPH=1
python tools/config3_profile.py 125000 > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --set full --import-source on --clock-control none -k regex:"wsi_(surfaces|step|settle)_kernel|wsi_mb_finalize" -s 12 -c 4 -f -o gpurun_out/${tag} \
    python tools/config3_profile.py 125000 > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_plain.log; tail -3 gpurun_out/${tag}_ncu.log
