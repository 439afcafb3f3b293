#!/bin/bash
# ncu recipes (B200_PROFILING.md) for this repo; run under gpurun from the repo root.  Outputs in gpurun_out/.
# usage: bash profiles/ncu_run.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
# 1. every launch of one eager step with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -s 57 -c 57 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1
# 2. full capture of the first conv_umma launches of a forward: fused stem (stem0+stem1), fused stem2+stem3, stage-0 convs
ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 0 -c 8 -o gpurun_out/conv_$TAG \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_conv_$TAG.log 2>&1
