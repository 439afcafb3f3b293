#!/bin/bash
# ncu recipes (B200_PROFILING.md) for this repo; run under gpurun from the repo root.  Outputs in gpurun_out/.
# usage: bash profiles/ncu_run.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
# 1. every launch of ONE eager step (forward + post-process) with its device time (cold-cache, serialised: compare SHARES)
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --ncu-step > gpurun_out/ncu_launches_$TAG.log 2>&1
# 2. full capture of the conv_umma launches of that step that matter: fused stem0+stem1, fused stem2+stem3, the stage-0 convs
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_umma -c 6 -o gpurun_out/conv_$TAG \
    python bench.py --ncu-step > gpurun_out/ncu_conv_$TAG.log 2>&1
