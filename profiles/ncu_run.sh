#!/bin/bash
# ncu recipes (B200_PROFILING.md) for this repo; run under gpurun from the repo root.  Outputs in gpurun_out/.
# usage: bash profiles/ncu_run.sh <tag>        then, here: python profiles/ncu_parse.py gpurun_out <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
# 1. every launch of ONE eager inference step (forward + post-process) with its device time (cold-cache, serialised: compare SHARES)
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --ncu-step > gpurun_out/ncu_launches_$TAG.log 2>&1
# 2. full capture of the first conv_umma launches of that step: fused stem0+stem1 (op 0), fused stem2+stem3 (op 1 = the dominant kernel), op 2, op 3
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_umma -c 4 -o gpurun_out/conv_$TAG \
    python bench.py --ncu-step > gpurun_out/ncu_conv_$TAG.log 2>&1
ncu -i gpurun_out/conv_$TAG.ncu-rep --page raw --csv --print-units base > gpurun_out/conv_${TAG}_raw.csv 2> /dev/null
# 3. the training step (WIDERFACE-L 640x640, 16 crops): launch list, and a full capture of the weight-gradient / norm-backward / head-backward kernels
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train_$TAG.csv \
    python bench.py --config WIDERFACE_L_train --ncu-step > gpurun_out/ncu_launches_train_$TAG.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"wgrad_umma|norm_bwd|head_final_bwd|bn_stats" -c 12 \
    -o gpurun_out/train_$TAG python bench.py --config WIDERFACE_L_train --ncu-step > gpurun_out/ncu_train_$TAG.log 2>&1
ncu -i gpurun_out/train_$TAG.ncu-rep --page raw --csv --print-units base > gpurun_out/train_${TAG}_raw.csv 2> /dev/null
