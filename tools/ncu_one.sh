#!/bin/bash
# usage: bash tools/ncu_one.sh <tag> <kernel regex> "<LBM_B200_TUNING>" nx ny dtype [steps coll turb]
# one full ncu capture (2 launches) of a step kernel, only after the same command exited 0 without ncu
TAG=$1; KRE=$2; export LBM_B200_TUNING="$3"; shift 3
python tools/prof_case.py "$@" > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KRE -s 2 -c 2 -o gpurun_out/${TAG} -f \
    python tools/prof_case.py "$@" > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu $TAG rc=$?"; cat gpurun_out/${TAG}_plain.log
