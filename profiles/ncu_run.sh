#!/bin/bash
# ncu recipes (B200_PROFILING.md) for this repo; run under gpurun from the repo root.  Outputs in gpurun_out/.
# usage: bash profiles/ncu_run.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
# 1. every launch of one eager step with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -s 55 -c 57 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1
# 2. full capture of the conv kernel instances (3 launches starting at the stem 1x1) and of the stem kernel
ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 0 -c 12 -o gpurun_out/conv_$TAG \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_conv_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stem0 -s 0 -c 1 -o gpurun_out/stem0_$TAG \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_stem_$TAG.log 2>&1
