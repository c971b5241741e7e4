#!/bin/bash
# round 2: ncu --set full of the FP64 row-GP kernel (fused, C4 in double)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r02z}
CMD="python bench.py --workload c4f64 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 200 $CMD > gpurun_out/${T}_plain.log 2>&1 && timeout 400 ncu --set full --clock-control none --import-source on -k regex:"RowGp64Kernel" -s 3 -c 1 -f -o gpurun_out/${T}_prof_rowgp64 $CMD > gpurun_out/${T}_ncu_full.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/${T}_ncu_full.log
