cd /root/repo
timeout 600 python -m pytest tests/test_gpu_batch.py tests/test_gpu_rowgp_tc.py tests/test_gpu_sensors.py -x -q 2>&1 | tail -3
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_batch.py -x -q -k "deterministic or multi_device or c4_shape" 2>&1 | tail -1; done
