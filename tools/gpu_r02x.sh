#!/bin/bash
# round 2: FP64 row-GP kernel (DMMA): parity + bench A/B against the generic kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r02x}
timeout 900 python -m pytest tests/test_gpu_batch.py tests/test_gpu_sensors.py -x -q -k "float64 or f64 or not_spd or limits or gate" 2>&1 | tail -25 > gpurun_out/${T}_tests.txt
tail -12 gpurun_out/${T}_tests.txt
for L in 0 1; do
if [ $L = 1 ]; then export ERL_GP_BATCH_LEGACY=1; fi
timeout 300 python bench.py --workload c4f64 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench_c4f64_legacy$L.json 2> gpurun_out/${T}_bench_c4f64_legacy$L.err
echo "legacy=$L"; cut -c1-220 gpurun_out/${T}_bench_c4f64_legacy$L.json; tail -2 gpurun_out/${T}_bench_c4f64_legacy$L.err
done
