#!/bin/bash
# Fusion kernel: plain timings, then one ncu --set full capture at 4096 queries x 3 x 100 candidates.
TAG=$1
mkdir -p gpurun_out
python tools/gpu_fuse.py > gpurun_out/fuse_$TAG.log 2>&1
echo "plain exit $?"; cat gpurun_out/fuse_$TAG.log
ncu --set full --clock-control none --import-source on -k regex:fuse_kernel -s 14 -c 1 -f -o gpurun_out/prof_fuse_$TAG \
    python tools/gpu_fuse.py 4096 100 100 > gpurun_out/ncu_full_fuse_$TAG.log 2>&1
echo "full capture exit $?"
