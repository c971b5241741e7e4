#!/bin/bash
# round 2: DMMA diagonal-block kernel in the dense Cholesky: parity + c1 / C5 Potrf / spgp timings, A/B against the generic kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_dense.py tests/test_gpu_full_size.py tests/test_gpu_noisy.py tests/test_pybind_module.py -x -q 2>&1 | tail -3
for L in 0 1; do
if [ $L = 1 ]; then export ERL_GP_DIAG_LEGACY=1; fi
echo "legacy=$L"
for W in c1 spgp; do timeout 200 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$W', round(d.get('ms_per_step'),3))"; done
timeout 300 python tools/bench_dense.py --n 16384 --t 4096 --dtype f64 2>&1 | tail -1 | cut -c1-160
timeout 300 python tools/bench_dense.py --n 1024 --t 8192 --dtype f64 2>&1 | tail -1 | cut -c1-160
done
