#!/bin/bash
# one GPU call (run twice in round 2; pass 1 chose the present defaults): A/B of the opt-in kernel variants at the bench sizes, then the whole validation with the
# winners switched on through the environment (what the defaults become afterwards): -m gpu suite, smoke(), the bench line, the ncu
# launch list of the bench command, and full captures of the group-by kernels (default + paired) and the pipeline kernels
mkdir -p gpurun_out
DEADLINE=$(( $(date +%s) + ${GPU_R2C_BUDGET_S:-780} ))
# run CMD... for at most LIMIT seconds, and not past the script's deadline (steps are ordered by priority; late ones may be skipped)
step() { local limit=$1; shift; local left=$(( DEADLINE - $(date +%s) )); if [ $left -lt 20 ]; then echo "skipped (deadline): $*" >> gpurun_out/c_skipped.txt; return 0; fi; [ $left -lt $limit ] && limit=$left; timeout $limit "$@"; }
step 240 python scripts/ab_variants.py 100 > gpurun_out/c_ab.json 2> gpurun_out/c_ab.err
eval "$(python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/c_ab.json"))
    print(f"export DFGPU_PIPE_VAR={d['q3']['winner']} DFGPU_AGG_PAIRED={d['c3']['winner_paired']} DFGPU_AGG_R4={d['c3']['winner_r4']}")
except Exception as e:
    print("export DFGPU_PIPE_VAR=11 DFGPU_AGG_PAIRED=1 DFGPU_AGG_R4=0")   # the library's defaults
PY
)"
echo "chosen: DFGPU_PIPE_VAR=$DFGPU_PIPE_VAR DFGPU_AGG_PAIRED=$DFGPU_AGG_PAIRED DFGPU_AGG_R4=$DFGPU_AGG_R4" | tee gpurun_out/c_chosen.txt
step 360 python -m pytest tests -x -q -m gpu --durations=12 2>&1 | tail -40 > gpurun_out/c_pytest.log
step 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c_smoke.log 2>&1
step 300 python bench.py > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err
# full captures of the kernels that changed (summarised on the box: the reports themselves are large to pull back), then the launch list
step 120 ncu --set full --clock-control none --import-source on -k regex:agg_update_pair -c 1 -f -o gpurun_out/r2c_agg python scripts/prof_c3_agg.py 1000000000 ${DFGPU_AGG_PAIRED} > gpurun_out/c_ncu_agg.log 2>&1
[ -f gpurun_out/r2c_agg.ncu-rep ] && python scripts/ncu_summary.py gpurun_out/r2c_agg.ncu-rep gpurun_out/r2c_agg_pair_kernel_268M_chunk.csv >> gpurun_out/c_ncu_agg.log 2>&1 && rm -f gpurun_out/r2c_agg.ncu-rep
step 120 ncu --set full --clock-control none --import-source on -k regex:pipe_kernel -c 3 -f -o gpurun_out/r2c_pipe python scripts/prof_q3_fused.py 100 1 > gpurun_out/c_ncu_pipe.log 2>&1
[ -f gpurun_out/r2c_pipe.ncu-rep ] && python scripts/ncu_summary.py gpurun_out/r2c_pipe.ncu-rep gpurun_out/r2c_pipe_kernels_sf100_final.csv >> gpurun_out/c_ncu_pipe.log 2>&1 && rm -f gpurun_out/r2c_pipe.ncu-rep
step 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2c_launch_list_bench.csv python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline --e2e-steps 0 > gpurun_out/c_launch_list.log 2>&1
cat gpurun_out/c_chosen.txt; tail -c 1500 gpurun_out/c_ab.json; tail -5 gpurun_out/c_ab.err; tail -6 gpurun_out/c_pytest.log; tail -2 gpurun_out/c_smoke.log; tail -3 gpurun_out/c_bench.err; tail -c 400 gpurun_out/c_bench.json; tail -2 gpurun_out/c_ncu_agg.log; tail -2 gpurun_out/c_ncu_pipe.log
