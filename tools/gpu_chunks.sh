#!/bin/bash
# A/B: chunks of the host-buffer pipeline (ERL_GP_BATCH_CHUNKS) on the C4 e2e step
cd "$(dirname "$0")/.."
for C in 4 6 8 10 12; do
ERL_GP_BATCH_CHUNKS=$C timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-tc-variant --no-other-workloads 2>/dev/null > /tmp/chunks_$C.json
python - <<PY
import json
d=json.loads([l for l in open('/tmp/chunks_$C.json') if l.startswith('{')][-1])
print('chunks', $C, round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3))
PY
done
