cd /root/repo
for N in 4 8; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02p_scale$N.json 2> gpurun_out/r02p_scale$N.err
echo "N=$N rc=$?"; tail -2 gpurun_out/r02p_scale$N.err
done
python - <<'PY'
import json
for N in (4,8):
    try:
        d=json.loads([l for l in open(f'gpurun_out/r02p_scale{N}.json') if l.startswith('{')][-1])
        print(N,{k:d[k] for k in ('value','ms_per_step')},'e2e',d['e2e']['ms_per_step'],d['e2e']['value']); s=d.get('strong'); print('  strong kernel',s['kernel'],'\n  strong e2e',s['e2e']['ms_per_step'],s['e2e']['value'],'\n  single-process',s.get('single_process_multi_device'))
    except Exception as e: print(N,'failed',e)
PY
timeout 300 python -m pytest tests/test_gpu_batch.py tests/test_gpu_cpp_dropin.py -q -k "multi_device or drop" 2>&1 | tail -2
