#!/bin/bash
# round 2: partition_on_hit_rays + large-GP path (n > 256): new tests first, then the whole GPU suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r02s}
timeout 900 python -m pytest tests/test_gpu_batch.py tests/test_gpu_sensors.py tests/test_gpu_cpp_dropin.py -x -q -k "large_n or limits or hit_rays or cpp_drop" 2>&1 | tail -30 > gpurun_out/${T}_new_tests.txt
tail -15 gpurun_out/${T}_new_tests.txt
timeout 1500 python -m pytest tests -x -q -m gpu --durations=8 2>&1 | tail -30 > gpurun_out/${T}_pytest_gpu.txt
tail -14 gpurun_out/${T}_pytest_gpu.txt
