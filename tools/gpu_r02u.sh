#!/bin/bash
# round 2, 8 GPUs: raw H2D ceiling of the box (k concurrent 205 MB uploads) + the scaling bench at N = 8
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r02u}
timeout 200 python tools/pcie_rate_multi.py 205 > gpurun_out/${T}_pcie_multi.json 2> gpurun_out/${T}_pcie_multi.err
cat gpurun_out/${T}_pcie_multi.json; tail -2 gpurun_out/${T}_pcie_multi.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/${T}_scale8.json 2> gpurun_out/${T}_scale8.err
echo "rc=$?"; tail -2 gpurun_out/${T}_scale8.err; cut -c1-300 gpurun_out/${T}_scale8.json
