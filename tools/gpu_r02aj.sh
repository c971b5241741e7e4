#!/bin/bash
# round 2: chunks of the host-buffer pipeline on two compute streams: parity (host path, multi-device) + e2e A/B
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_batch.py tests/test_gpu_rowgp_tc.py tests/test_gpu_full_size.py -x -q -k "not c5 and not spgp" 2>&1 | tail -3
for O in 0 1; do
if [ $O = 1 ]; then export ERL_GP_BATCH_ONE_COMPUTE_STREAM=1; fi
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-tc-variant --no-other-workloads 2>/dev/null > /tmp/cs_$O.json
python - <<PY
import json
d=json.loads([l for l in open('/tmp/cs_$O.json') if l.startswith('{')][-1])
print('one_stream', $O, round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3))
PY
done
