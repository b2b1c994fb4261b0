#!/bin/bash
# Select kernels: plain timing of every shape, then one ncu --set full capture of the slice kernel at 256 x 1M.
# usage: tools/gpu_select_prof.sh <tag>
TAG=$1
mkdir -p gpurun_out
python tools/gpu_select.py > gpurun_out/select_$TAG.log 2>&1
echo "plain exit $?"; cat gpurun_out/select_$TAG.log
python tools/gpu_select.py 256 1000000 100 > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:topk_s -s 3 -c 1 -f -o gpurun_out/prof_select_$TAG \
    python tools/gpu_select.py 256 1000000 100 > gpurun_out/ncu_full_select_$TAG.log 2>&1
echo "full capture exit $?"
