#!/bin/bash
# round 2: tc kernel iteration: tests + bench + timing variant
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r02d}
( timeout 300 python -m pytest tests/test_gpu_rowgp_tc.py -x -q 2>&1 | tail -25 > gpurun_out/${T}_tc_tests.txt; echo "exit ${PIPESTATUS[0]}" >> gpurun_out/${T}_tc_tests.txt )
tail -6 gpurun_out/${T}_tc_tests.txt
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench_c4.json 2> gpurun_out/${T}_bench_c4.err
cut -c1-330 gpurun_out/${T}_bench_c4.json; tail -3 gpurun_out/${T}_bench_c4.err
ERL_GP_B200_LIB=$PWD/erl_gaussian_process_b200/lib/liberl_gp_b200_timing.so timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${T}_timing.txt 2>&1
grep "tc timing" gpurun_out/${T}_timing.txt | tail -5
timeout 300 ncu --metrics sm__icc_request_hit_rate.pct,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,gpu__time_duration.sum --clock-control none -k regex:"RowGpTcKernel" -s 3 -c 1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep -E "icc|inst_executed|issue_active|no_instruction|duration" 
