#!/bin/bash
# round 2, call A: tcgen05 probe, the whole -m gpu suite, bench lines of every BASELINE config
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_gpu.txt 2>&1
( timeout 120 tools/tcgen05_probe > gpurun_out/r02a_probe.txt 2>&1; echo "probe exit $?" >> gpurun_out/r02a_probe.txt )
( timeout 1500 python -m pytest tests -x -q -m gpu --durations=15 > gpurun_out/r02a_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r02a_pytest.txt )
for w in c1 c2 c3 c3n256 spgp; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/r02a_bench_$w.json 2> gpurun_out/r02a_bench_$w.err
done
timeout 600 python bench.py --workload c5 --steps 2 --warmup 3 > gpurun_out/r02a_bench_c5.json 2> gpurun_out/r02a_bench_c5.err
timeout 300 python bench.py --steps 20 --warmup 3 --e2e-variants > gpurun_out/r02a_bench_c4.json 2> gpurun_out/r02a_bench_c4.err
tail -3 gpurun_out/r02a_probe.txt gpurun_out/r02a_pytest.txt
