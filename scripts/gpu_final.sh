#!/bin/bash
# end-of-round validation on one GPU: the full -m gpu suite, smoke(), the bench line, and the profile captures that profiles/ quotes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/z_pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/z_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/z_bench_ref.json 2> gpurun_out/z_bench_ref.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:radix_ -c 8 -f -o gpurun_out/r2b_radix python scripts/prof_c2_radix.py > gpurun_out/z_ncu_radix.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2b_launch_list_bench.csv python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline --e2e-steps 0 > gpurun_out/z_launch_list.log 2>&1
tail -4 gpurun_out/z_pytest_all.log; tail -2 gpurun_out/z_smoke.log; tail -2 gpurun_out/z_bench.err; tail -c 600 gpurun_out/z_bench_ref.json; tail -2 gpurun_out/z_ncu_radix.log
