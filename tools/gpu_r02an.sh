#!/bin/bash
# TMA-fed FP64 GEMM (bulk copies + mbarrier pipeline): dense parity tests, then A/B of the C5 train and the spgp workload
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_full_size.py tests/test_gpu_noisy.py -m gpu -x -q > gpurun_out/r02an_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r02an_tests.log
tail -3 gpurun_out/r02an_tests.log
for tma in 0 1 0 1; do
  echo "ERL_GP_DENSE_TMA=$tma"
  ERL_GP_DENSE_TMA=$tma timeout 300 python tools/bench_dense.py --n 16384 --t 16384 --reps 3 2>&1 | tail -1 | cut -c1-400
done
for tma in 0 1; do
  ERL_GP_DENSE_TMA=$tma timeout 300 python bench.py --workload spgp --steps 10 --warmup 3 --no-cpu-baseline --no-other-workloads 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('spgp tma=$tma', d['ms_per_step'])"
done
