#!/bin/bash
# round 2: the whole GPU suite after PartitionOnHitRays / large-GP path / NoisyInputGaussianProcess / measured tolerances
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r02v}
timeout 2400 python -m pytest tests -x -q -m gpu --durations=8 -s 2>&1 | grep -v "^$" | tail -60 > gpurun_out/${T}_pytest_gpu.txt
tail -40 gpurun_out/${T}_pytest_gpu.txt
