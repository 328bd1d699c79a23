#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_join_wide_keys.py tests/test_gpu_join.py tests/test_gpu_exec_api.py -q --tb=short 2>&1 | tail -60 > gpurun_out/f_pytest.log
grep -v "^$" gpurun_out/f_pytest.log | tail -45
