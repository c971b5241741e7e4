#!/bin/bash
# round 2: role rotation of the mma.sync row-GP kernel (pivot warps on different SM sub-partitions): A/B + per-SMSP balance
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r02r}
for R in 0 1; do
ERL_GP_ROWGP_ROTATE=$R timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-tc-variant > gpurun_out/${T}_bench_c4_rot$R.json 2> gpurun_out/${T}_bench_c4_rot$R.err
echo "rotate=$R"; cut -c1-200 gpurun_out/${T}_bench_c4_rot$R.json; tail -2 gpurun_out/${T}_bench_c4_rot$R.err
done
for W in n192 n256 c3; do for R in 0 1; do
echo "$W rotate=$R"; ERL_GP_ROWGP_ROTATE=$R timeout 200 python bench.py --workload $W --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | cut -c1-160
done; done
( ERL_GP_ROWGP_ROTATE=1 timeout 400 python -m pytest tests/test_gpu_batch.py -x -q 2>&1 | tail -5 )
for R in 0 1; do
echo "ncu rotate=$R"
ERL_GP_ROWGP_ROTATE=$R timeout 300 ncu --metrics smsp__inst_executed.min,smsp__inst_executed.max,smsp__inst_executed.avg,smsp__issue_active.min,smsp__issue_active.max,smsp__issue_active.avg,smsp__cycles_active.avg,gpu__time_duration.sum --clock-control none -k regex:"RowGpKernel" -s 3 -c 1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-tc-variant 2>&1 | grep -E "inst_executed|issue_active|cycles_active|duration" | tee gpurun_out/${T}_ncu_smsp_rot$R.txt
done
