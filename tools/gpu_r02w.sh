#!/bin/bash
# round 2, 8 GPUs: replicate / test_multi over real peers (test + C5 from one process)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r02w}
timeout 600 python -m pytest tests/test_gpu_dense.py tests/test_gpu_batch.py -x -q -k "replicate or multi_device" 2>&1 | tail -6
timeout 600 python tools/bench_c5_multi.py 16384 1000000 > gpurun_out/${T}_c5_multi.json 2> gpurun_out/${T}_c5_multi.err
cat gpurun_out/${T}_c5_multi.json; tail -3 gpurun_out/${T}_c5_multi.err
