#!/bin/bash
# round 2: FP64 row-GP kernel iteration: parity + fused / train / predict timings
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r02y}
timeout 900 python -m pytest tests/test_gpu_batch.py tests/test_gpu_sensors.py -x -q -k "float64 or f64 or not_spd or limits or gate" 2>&1 | tail -4
for ph in fused train predict; do
timeout 200 python bench.py --workload c4f64 --phase $ph --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$ph', round(d.get('ms_per_step'),3))"
done
