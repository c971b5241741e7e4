#!/bin/bash
# final state of a round: smoke, the whole GPU suite, the default bench line (with other_workloads), the reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-final}
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/${T}_pytest_gpu.txt; tail -2 gpurun_out/${T}_pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/${T}_bench_c4.json 2> gpurun_out/${T}_bench_c4.err; cut -c1-260 gpurun_out/${T}_bench_c4.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; cut -c1-200 gpurun_out/${T}_bench_ref.json
