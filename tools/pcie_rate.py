"""Raw pinned-host <-> device copy bandwidth of the box (the denominator of the e2e number: bench.py uploads 205 MB per step)."""
import torch, time
dev=torch.device('cuda:0')
for mb in (3, 10, 26, 205):
    n=mb*1024*1024
    h=torch.empty(n,dtype=torch.uint8).pin_memory(); d=torch.empty(n,dtype=torch.uint8,device=dev)
    torch.cuda.synchronize()
    for _ in range(3): d.copy_(h,non_blocking=True)
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(10): d.copy_(h,non_blocking=True)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/10
    h2=torch.empty(n,dtype=torch.uint8).pin_memory()
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(10): h2.copy_(d,non_blocking=True)
    torch.cuda.synchronize(); dt2=(time.perf_counter()-t0)/10
    print(f"{mb} MB: H2D {n/dt/1e9:.1f} GB/s  D2H {n/dt2/1e9:.1f} GB/s")
