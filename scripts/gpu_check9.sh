#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_exec_api.py tests/test_gpu_dictionary.py tests/test_c_consumer.py -q --tb=short 2>&1 | tail -30 > gpurun_out/k_pytest.log
grep -v "^$" gpurun_out/k_pytest.log | tail -25
