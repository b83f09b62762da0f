#!/bin/bash
# ncu recipe of this repo (run under gpurun): launch list of a bench run + full capture of the hot kernels.
# usage: profiles/run_ncu.sh <tag>
set -x
TAG=${1:-r1}
mkdir -p gpurun_out
# every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes); skip the warm-up steps
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 450 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-profile \
    > gpurun_out/launches_${TAG}.log 2>&1
# the hot kernels, full sections (one launch each of the triplet k/v, knn k/v and the gemm)
ncu --set full --clock-control none --import-source on -k regex:'trip_kernel|knn_attn_k_kernel|knn_attn_v_node_kernel|gemm128_kernel' \
    -s 40 -c 8 -o gpurun_out/prof_${TAG} python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-profile \
    > gpurun_out/prof_${TAG}.log 2>&1
ls -la gpurun_out
