#!/bin/bash
# round 2: one-pass diagonal block + inverse in the generic Cholesky (DiagFactorKernel of the dense path, generic batched kernels)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for W in c1 spgp; do timeout 200 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$W', round(d.get('ms_per_step'),3), d.get('phases'))"; done
timeout 300 python tools/bench_dense.py --n 16384 --t 4096 --dtype f64 2>&1 | tail -2 | cut -c1-400
ERL_GP_BATCH_LEGACY=1 timeout 200 python bench.py --workload c4f64 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('c4f64 generic kernel', round(d.get('ms_per_step'),3))"
