#!/bin/bash
# round-1b ncu recipe (run under gpurun): launch list of one bench run + full capture of the tensor-core kernels.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 450 --csv \
    --log-file gpurun_out/launches_r1b.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-profile \
    > gpurun_out/launches_r1b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'knn_tc_kernel|trip_tc_kernel|gemm128_tc_kernel' \
    -s 30 -c 10 -o gpurun_out/prof_r1b -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-profile \
    > gpurun_out/prof_r1b.log 2>&1
ncu -i gpurun_out/prof_r1b.ncu-rep --page raw --csv > gpurun_out/raw_r1b.csv 2>/dev/null
ls -la gpurun_out
