#!/bin/bash
# round 2, final state: smoke, the whole GPU suite, default bench line, ncu launch list + full capture of the default kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r02ac}
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/${T}_pytest_gpu.txt; tail -2 gpurun_out/${T}_pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/${T}_bench_c4.json 2> gpurun_out/${T}_bench_c4.err; cut -c1-260 gpurun_out/${T}_bench_c4.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; cut -c1-300 gpurun_out/${T}_bench_ref.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-tc-variant --no-other-workloads"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/${T}_ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"RowGpKernel" -s 3 -c 1 -f -o gpurun_out/${T}_prof_rowgp $CMD > gpurun_out/${T}_ncu_full.log 2>&1
echo "ncu full rc=$?"
