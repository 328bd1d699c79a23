#!/bin/bash
# one GPU call, many answers: parity tests, the Q3 timing A/B, an ncu capture of the pipeline kernels, the full bench line
mkdir -p gpurun_out
python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_q3_device_pipeline.py -x -q 2>&1 | tail -15 > gpurun_out/s_pytest_pipeline.log
python scripts/q3_fused_timing.py 100 > gpurun_out/s_q3_timing.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pipe_kernel -c 4 -o gpurun_out/s_pipe python scripts/prof_q3_fused.py 30 1 > gpurun_out/s_ncu.log 2>&1
for b in 0; do for h in 15 7 3 11; do echo "blocks_per_sm=$b hints=$h"; DFGPU_PIPE_BLOCKS_PER_SM=$b DFGPU_PIPE_HINTS=$h python scripts/q3_fused_timing.py 100 2>&1 | grep -A4 '"kernel_ms"' | head -5; done; done > gpurun_out/s_sweep.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/s_pytest_all.log
tail -4 gpurun_out/s_pytest_pipeline.log; cat gpurun_out/s_sweep.log; grep -A5 '"kernel_ms"' gpurun_out/s_q3_timing.log | head -14; tail -c 1500 gpurun_out/s_bench.json; tail -5 gpurun_out/s_bench.err; tail -4 gpurun_out/s_pytest_all.log
