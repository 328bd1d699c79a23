#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_q3_device_pipeline.py tests/test_gpu_decimal.py tests/test_gpu_agg.py -q --tb=short 2>&1 | tail -40 > gpurun_out/e_pytest.log
timeout 900 python bench.py > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err
tail -12 gpurun_out/e_pytest.log; tail -3 gpurun_out/e_bench.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/e_bench.json') if l.startswith('{')][0])
print(d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'])
print(d['roofline']['secondary'].get('C4_q3_decimal128_money'))
PY
