#!/bin/bash
# round 2: fast FP64 exp / sqrt for every covariance entry: whole GPU suite + the double-precision workloads
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for W in c1 c4f64 spgp; do timeout 200 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$W', round(d.get('ms_per_step'),3))"; done
timeout 300 python tools/bench_dense.py --n 1024 --t 8192 --dtype f64 2>&1 | tail -1 | cut -c1-200
timeout 300 python tools/bench_dense.py --n 16384 --t 65536 --dtype f64 2>&1 | tail -1 | cut -c1-260
