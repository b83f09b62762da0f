#!/bin/bash
# full ncu capture of selected kernels: profiles/ncu_one.sh <tag> <kernel regex> <skip> <count>
TAG=$1; RE=$2; SKIP=${3:-20}; CNT=${4:-6}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c $CNT -o gpurun_out/prof_${TAG} -f \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-profile > gpurun_out/prof_${TAG}.log 2>&1
ncu -i gpurun_out/prof_${TAG}.ncu-rep --page raw --csv > gpurun_out/raw_${TAG}.csv 2>/dev/null
ncu -i gpurun_out/prof_${TAG}.ncu-rep --page source --csv > gpurun_out/src_${TAG}.csv 2>/dev/null
