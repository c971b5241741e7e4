cd /root/repo
timeout 300 python -m pytest tests/test_gpu_rowgp_tc.py -q 2>&1 | tail -15
timeout 300 python -m pytest tests/test_gpu_batch.py -x -q 2>&1 | tail -5
