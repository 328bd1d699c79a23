#!/bin/bash
for b in 1; do for pct in 125 150 200 250; do echo "JOIN_BUCKET=$b CAP_PCT=$pct"; DFGPU_JOIN_BUCKET=$b DFGPU_JOIN_CAP_PCT=$pct python scripts/phase_timing.py sparse 2>&1 | tail -2 | head -1; done; done
echo "JOIN_BUCKET=0 CAP_PCT=250"; DFGPU_JOIN_BUCKET=0 DFGPU_JOIN_CAP_PCT=250 python scripts/phase_timing.py sparse 2>&1 | tail -2 | head -1
python -m pytest tests/test_gpu_join.py tests/test_gpu_exec_api.py tests/test_gpu_tpch_q3.py -x -q -m gpu --timeout 300 2>&1 | tail -2
