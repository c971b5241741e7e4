#!/bin/bash
# round 2, call B: first runs of the tcgen05 fused kernel (bounded by timeouts: a protocol error traps / is killed)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_rowgp_tc.py -x -q 2>&1 | tail -40 > gpurun_out/r02b_tc_tests.txt; echo "exit ${PIPESTATUS[0]}" >> gpurun_out/r02b_tc_tests.txt )
( timeout 300 python -m pytest tests/test_gpu_batch.py -x -q 2>&1 | tail -30 > gpurun_out/r02b_batch_tests.txt; echo "exit ${PIPESTATUS[0]}" >> gpurun_out/r02b_batch_tests.txt )
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench_c4_tc.json 2> gpurun_out/r02b_bench_c4_tc.err
ERL_GP_ROWGP_TC=0 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02b_bench_c4_old.json 2> gpurun_out/r02b_bench_c4_old.err
tail -5 gpurun_out/r02b_tc_tests.txt gpurun_out/r02b_batch_tests.txt
cut -c1-400 gpurun_out/r02b_bench_c4_tc.json gpurun_out/r02b_bench_c4_old.json
tail -3 gpurun_out/r02b_bench_c4_tc.err
