#!/bin/bash
# Round-2 profile pass (one GPU): launch list + DRAM traffic of the bench command, and one
# `ncu --set full` capture per solve kernel at a reduced iteration budget (a --set full replay of the
# full-length kernels would take tens of minutes).  Outputs land in gpurun_out/.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --latency-solves 0 \
    > gpurun_out/r2_bench_under_ncu.log 2>&1
M=sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active
for cfg in "warehouse_b65536x8 2368 100 2 default" "warehouse_b4096_ndyn40 4096 50 2 ndyn40" "dense_crowd_n40 592 20 2 dense"; do
  set -- $cfg
  ncu --set full --import-source on --clock-control none -k regex:solve_kernel -c 1 -f -o gpurun_out/r2_full_$5 \
      python scripts/profile_config.py $1 $2 $3 $4 1 > gpurun_out/r2_full_$5.log 2>&1
done
