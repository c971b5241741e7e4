#!/bin/bash
# One gpurun call: tests -> bench -> ncu launch list -> ncu full capture of the top kernel.
# usage: tools/gpu_bench_profile.sh <tag>
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench rc=$?"
cat gpurun_out/bench_${TAG}.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_list_${TAG}.log 2>&1
echo "ncu list rc=$?"
$CMD > gpurun_out/plain2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"RowGpKernel|BatchedGpKernel" -s 3 -c 1 -f -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out | tail -12
