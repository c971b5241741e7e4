#!/bin/bash
# round 2: rcp-only pivot chain in the f32 row-GP kernels + the FP64 row-GP kernel: whole GPU suite, c4 / c4f64 bench lines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r02aa}
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/${T}_pytest_gpu.txt
tail -4 gpurun_out/${T}_pytest_gpu.txt
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/${T}_bench_c4.json 2> gpurun_out/${T}_bench_c4.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/${T}_bench_c4.json') if l.startswith('{')][-1])
print('c4', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'tc', d.get('tcgen05_variant',{}).get('ms_per_step'), 'frac', d['roofline']['frac'])
PY
timeout 400 python bench.py --workload c4f64 --steps 10 --warmup 3 > gpurun_out/${T}_bench_c4f64.json 2> gpurun_out/${T}_bench_c4f64.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/${T}_bench_c4f64.json') if l.startswith('{')][-1])
print('c4f64', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['roofline'].get('fp64_tensor_pipe'), 'cpu', d['cpu_baseline']['value'])
PY
