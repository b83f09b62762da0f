#!/bin/bash
# round-1d ncu recipe (run under gpurun): launch list of one bench run + full captures of the tensor-core kernels.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 340 -c 360 --csv \
    --log-file gpurun_out/launches_r1g.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-profile \
    > gpurun_out/launches_r1g.log 2>&1
for spec in "knn:knn_tc_kernel:0:4" "trip:trip_tc_kernel:0:2" "bond:bond_tc_kernel:0:4" "gemm:gemm128_tc_kernel:0:12"; do
  IFS=: read tag re skip cnt <<< "$spec"
  ncu --set full --clock-control none --import-source on -k regex:"$re" -s $skip -c $cnt -o gpurun_out/prof_r1g_$tag -f \
      python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-profile > gpurun_out/prof_r1g_$tag.log 2>&1
  ncu -i gpurun_out/prof_r1g_$tag.ncu-rep --page raw --csv > gpurun_out/raw_r1g_$tag.csv 2>/dev/null
done
ls -la gpurun_out | tail -12
