cd /root/repo
for N in 4; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02q_scale$N.json 2> gpurun_out/r02q_scale$N.err
echo "N=$N rc=$?"; tail -2 gpurun_out/r02q_scale$N.err
done
python - <<'PY'
import json
for N in (4,):
    d=json.loads([l for l in open(f'gpurun_out/r02q_scale{N}.json') if l.startswith('{')][-1])
    s=d.get('strong'); print('  strong e2e',s['e2e']['ms_per_step'],'\n  single-process',s.get('single_process_multi_device'))
PY
