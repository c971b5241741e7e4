#!/bin/bash
# round 2, call C: per-phase cycle counters of the tcgen05 fused kernel (timing variant) + ncu full capture of the product kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ERL_GP_B200_LIB=$PWD/erl_gaussian_process_b200/lib/liberl_gp_b200_timing.so timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02c_timing.txt 2>&1
grep "tc timing" gpurun_out/r02c_timing.txt | tail -6
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"RowGpTcKernel" -s 3 -c 1 -f -o gpurun_out/r02c_prof_tc $CMD > gpurun_out/r02c_ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out | tail -5
