#!/bin/bash
# round-2 profile captures (one GPU): launch list of the bench command + full captures of the dominant kernels
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launch_list_bench.csv python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline --e2e-steps 0 > gpurun_out/r2_launch_list_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pipe_kernel -c 3 -o gpurun_out/r2_pipe_sf100 python scripts/prof_q3_fused.py 100 1 > gpurun_out/r2_pipe_sf100.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:radix_ -c 8 -o gpurun_out/r2_radix python scripts/prof_c2_radix.py > gpurun_out/r2_radix.log 2>&1
tail -2 gpurun_out/r2_launch_list_bench.log; tail -2 gpurun_out/r2_pipe_sf100.log; tail -2 gpurun_out/r2_radix.log
