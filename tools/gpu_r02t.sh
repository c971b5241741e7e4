#!/bin/bash
# round 2: NoisyInputGaussianProcess parity, hit-ray partitions, c3n484 bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r02t}
timeout 900 python -m pytest tests/test_gpu_noisy.py tests/test_gpu_sensors.py -x -q -k "noisy or hit_rays" 2>&1 | tail -30 > gpurun_out/${T}_new_tests.txt
tail -25 gpurun_out/${T}_new_tests.txt
timeout 600 python bench.py --workload c3n484 --steps 10 --warmup 3 > gpurun_out/${T}_bench_c3n484.json 2> gpurun_out/${T}_bench_c3n484.err
cut -c1-400 gpurun_out/${T}_bench_c3n484.json; tail -3 gpurun_out/${T}_bench_c3n484.err
