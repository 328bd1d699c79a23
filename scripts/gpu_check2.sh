#!/bin/bash
# targeted GPU check: decimal + dictionary tests, pipeline + join tests, Q3 fused timing, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decimal.py tests/test_gpu_dictionary.py -q --tb=short 2>&1 | tail -150 > gpurun_out/d_pytest_decimal.log
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_join.py tests/test_gpu_filter.py tests/test_gpu_agg.py tests/test_gpu_exec_api.py -x -q 2>&1 | tail -8 > gpurun_out/d_pytest_rest.log
python scripts/q3_fused_timing.py 100 > gpurun_out/d_q3_timing.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err
grep -v "^$" gpurun_out/d_pytest_decimal.log | tail -60; tail -4 gpurun_out/d_pytest_rest.log; grep -A6 '"kernel_ms"' gpurun_out/d_q3_timing.log | head -8
