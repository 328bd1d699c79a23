#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_join.py tests/test_gpu_join_wide_keys.py -q --tb=short 2>&1 | tail -12 > gpurun_out/h_pytest.log
timeout 900 python bench.py --no-cpu-baseline --e2e-steps 1 > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err
tail -5 gpurun_out/h_pytest.log; tail -3 gpurun_out/h_bench.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/h_bench.json') if l.startswith('{')][0])
print(d['ms_per_step'], d['roofline']['kernel_ms'])
for k,v in d['roofline']['secondary'].items():
    print(k, round(v['ms_per_step'],3), round(v.get('frac') or 0,3))
PY
