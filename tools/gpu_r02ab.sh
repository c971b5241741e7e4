#!/bin/bash
# round 2: look-ahead factorisation of the FP32 row-GP kernel: parity + C4 timings (fused / train)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r02ab}
timeout 900 python -m pytest tests/test_gpu_batch.py tests/test_gpu_sensors.py tests/test_gpu_full_size.py -x -q -k "not c5 and not spgp and not noisy" 2>&1 | tail -5
for ph in fused train predict; do
timeout 200 python bench.py --phase $ph --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-tc-variant 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$ph', round(d.get('ms_per_step'),3))"
done
for W in c3 c3n256 c2; do timeout 200 python bench.py --workload $W --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$W', round(d.get('ms_per_step'),3))"; done
