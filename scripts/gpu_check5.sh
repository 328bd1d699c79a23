#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decimal.py tests/test_gpu_dictionary.py tests/test_gpu_filter.py tests/test_gpu_pipeline.py tests/test_gpu_q3_device_pipeline.py tests/test_short_circuit.py -q --tb=short 2>&1 | tail -12 > gpurun_out/g_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/g_smoke.log 2>&1
python scripts/q3_fused_timing.py 100 > gpurun_out/g_q3_timing.log 2>&1
tail -5 gpurun_out/g_pytest.log; tail -3 gpurun_out/g_smoke.log; grep -A6 '"kernel_ms"' gpurun_out/g_q3_timing.log | head -8; grep '"ms"' gpurun_out/g_q3_timing.log | head -2
