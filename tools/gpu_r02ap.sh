#!/bin/bash
# ncu evidence for the TMA-fed FP64 GEMM: the longest rank-512 trailing update of Potrf n = 16384, full capture, next to the
# register-staged kernel on the same launch
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python tools/bench_dense.py --n 16384 --t 128 --reps 1"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:GemmKernelDmmaTma -c 80 --csv --log-file gpurun_out/r02ap_tma_list.csv $CMD > gpurun_out/r02ap_list.log 2>&1
echo "list rc=$?"
IDX=$(python - <<'PY'
import csv
lines=open('gpurun_out/r02ap_tma_list.csv').read().splitlines()
hi=[i for i,l in enumerate(lines) if l.startswith('"ID"')][0]
rows=list(csv.reader(lines[hi:]))[1:]
d=[float(r[-1]) for r in rows]
print(max(range(len(d)), key=lambda i: d[i]))
PY
)
echo "longest TMA GEMM launch: index $IDX"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:GemmKernelDmmaTma -s $IDX -c 1 -f -o gpurun_out/r02ap_gemm_tma $CMD > gpurun_out/r02ap_full_tma.log 2>&1
echo "full tma rc=$?"
ERL_GP_DENSE_TMA=0 timeout 300 ncu --set full --clock-control none -k regex:GemmKernelDmma512 -s $IDX -c 1 -f -o gpurun_out/r02ap_gemm_regstaged $CMD > gpurun_out/r02ap_full_reg.log 2>&1
echo "full reg rc=$?"
