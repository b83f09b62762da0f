#!/bin/bash
# final round-2 measurement session (run under gpurun, 1 GPU): profiles/r2f_measure.sh
# benches of every single-GPU config, ligand-size and cfg-5 sweeps, launch list, ncu full captures of the dominant kernels
# (condensed with summarize_raw.py), compute-sanitizer on the final kernels.
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench_cfg2.json 2> gpurun_out/r2f_bench_cfg2.err
python bench.py --workload cfg3 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2f_bench_cfg3.json 2> gpurun_out/r2f_bench_cfg3.err
python bench.py --workload cfg1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2f_bench_cfg1.json 2> gpurun_out/r2f_bench_cfg1.err
for w in lig40 lig50 lig64; do python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2f_bench_$w.json 2> gpurun_out/r2f_bench_$w.err; done
python profiles/sweep_cfg5.py --steps 100 > gpurun_out/r2f_sweep_cfg5.jsonl 2> gpurun_out/r2f_sweep_cfg5.err
# every launch of 2 steps with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -s 260 -c 260 --csv --log-file gpurun_out/r2f_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-profile > gpurun_out/r2f_launches.log 2>&1
for spec in "knn:knn_tc:0:3" "trip:trip3_kernel:0:2" "bond:bond_tc:0:2" "gemm:gemm128_tc_kernel:1:7" "graph:knn_merge_kernel|knn_kernel|graph_levels_kernel|graph_lists_kernel|edge_weight_kernel|trip_prep_kernel:2:6"; do
  IFS=: read tag re skip cnt <<< "$spec"
  ncu --set full --clock-control none --import-source on -k regex:"$re" -s $skip -c $cnt -o gpurun_out/prof_r2f_$tag -f \
      python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-profile > gpurun_out/prof_r2f_$tag.log 2>&1
  ncu -i gpurun_out/prof_r2f_$tag.ncu-rep --page raw --csv > gpurun_out/raw_r2f_$tag.csv 2>/dev/null
  python profiles/summarize_raw.py gpurun_out/raw_r2f_$tag.csv > gpurun_out/r2f_ncu_full_$tag.csv
  rm -f gpurun_out/prof_r2f_$tag.ncu-rep
done
bash profiles/sanitize.sh
for t in memcheck racecheck; do cp gpurun_out/sanitize_$t.log gpurun_out/r2f_sanitize_$t.log; done
ls -la gpurun_out | tail -40
