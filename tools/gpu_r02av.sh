#!/bin/bash
# FP64 row-GP kernel, n <= 192 instance: parity tests, then A/B against the generic kernel on the n192f64 diagnostic workload
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batch.py tests/test_gpu_sensors.py tests/test_gpu_full_size.py tests/test_gpu_cpp_dropin.py -m gpu -x -q > gpurun_out/r02av_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r02av_tests.log
tail -4 gpurun_out/r02av_tests.log
for legacy in 1 0; do
  if [ $legacy = 1 ]; then export ERL_GP_BATCH_LEGACY_LARGE=1; else unset ERL_GP_BATCH_LEGACY_LARGE; fi
  timeout 300 python bench.py --workload n192f64 --steps 5 --warmup 3 --no-cpu-baseline --no-other-workloads --no-tc-variant 2>/dev/null | tail -1 > gpurun_out/r02av_n192f64_legacy$legacy.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02av_n192f64_legacy$legacy.json").read())
print("legacy_large=$legacy", d["ms_per_step"], d.get("e2e",{}).get("ms_per_step"))
PY
done
