#!/bin/bash
# TMA write-back of L in the FP64 row-GP kernel: parity tests, then A/B on the c4f64 workload
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batch.py tests/test_gpu_sensors.py tests/test_gpu_full_size.py tests/test_gpu_dense.py tests/test_gpu_cpp_dropin.py -m gpu -x -q > gpurun_out/r02am_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r02am_tests.log
tail -3 gpurun_out/r02am_tests.log
for wb in 0 1 0 1; do
  ERL_GP_ROWGP_TMA_WB=$wb timeout 300 python bench.py --workload c4f64 --steps 8 --warmup 3 --no-cpu-baseline --no-other-workloads 2>/dev/null | tail -1 > gpurun_out/r02am_bench_wb$wb.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02am_bench_wb$wb.json").read())
print("wb=$wb", d["ms_per_step"], d.get("e2e",{}).get("value"), d["roofline"]["frac"])
PY
done
