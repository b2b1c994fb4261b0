#!/bin/bash
# ncu launch list + one full capture of the dominant kernel, each preceded by the same command run plain.
# usage: tools/gpu_profile.sh <tag> <kernel-regex> <bench args...>
TAG=$1; KREGEX=$2; shift 2
mkdir -p gpurun_out
python bench.py "$@" --no-cpu-baseline > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:lrag:: -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py "$@" --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "launch list exit $?"
python bench.py "$@" --no-cpu-baseline > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s ${NCU_SKIP:-3} -c ${NCU_COUNT:-1} -f -o gpurun_out/prof_$TAG \
    python bench.py "$@" --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture exit $?"
