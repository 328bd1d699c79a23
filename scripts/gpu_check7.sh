#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_agg_wide_keys.py tests/test_gpu_agg.py tests/test_gpu_decimal.py tests/test_gpu_tpch_q3.py tests/test_gpu_exec_api.py -q --tb=short 2>&1 | tail -40 > gpurun_out/i_pytest.log
grep -v "^$" gpurun_out/i_pytest.log | tail -30
