#!/bin/bash
# full ncu capture (source-level stall sampling) of the first two conv_umma launches of a forward: fused stem0+stem1, fused stem2+stem3
TAG=${1:-r01d}
ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 0 -c 2 -o gpurun_out/conv_$TAG python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_conv_$TAG.log 2>&1
