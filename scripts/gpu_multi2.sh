#!/bin/bash
# 2-GPU sanity: the N > 1 tests and the bench line at N = 2 (driver launch line)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q --tb=short 2>&1 | tail -8 > gpurun_out/n2_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err
tail -4 gpurun_out/n2_pytest.log; tail -3 gpurun_out/n2_bench.err; tail -c 900 gpurun_out/n2_bench.json
