cd /root/repo
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/r02n_bench_c4.json 2> gpurun_out/r02n_bench_c4.err; echo rc=$?
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02n_bench_c4.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['roofline']['kernel'], d['roofline']['frac']); print('e2e',d['e2e']['ms_per_step']); print('tc',d.get('tcgen05_variant')); print('cpu',d.get('cpu_baseline'))
PY
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
