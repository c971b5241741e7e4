"""Ceiling of the host-buffer (e2e) step at N GPUs: N concurrent pinned-host -> device copies of 205 MB (the upload of one C4 step per
rank) from ONE process, one stream per device, timed with CUDA events on every device (max over devices) for k = 1, 2, 4, 8 active
devices.  If the aggregate stops growing with k, the limit is the host side (one NUMA node's DRAM / root complex feeding all links),
not the library's copy pipeline.  Prints one JSON line."""
import json
import sys

import torch

n_dev = torch.cuda.device_count()
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 205
nbytes = mb * 1024 * 1024
host = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(n_dev)]
dev = [torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{i}") for i in range(n_dev)]
back = [torch.empty(nbytes // 4, dtype=torch.uint8).pin_memory() for _ in range(n_dev)]  # 58 MB of results per step: ~ a quarter
streams = [torch.cuda.Stream(device=i) for i in range(n_dev)]
out = {"mb_per_device": mb, "devices": n_dev, "h2d": {}, "h2d_plus_d2h": {}}
for both in (False, True):
    for k in [v for v in (1, 2, 4, 8) if v <= n_dev]:
        reps = 10
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
        for warm in (True, False):
            for i in range(k):
                with torch.cuda.device(i), torch.cuda.stream(streams[i]):
                    if not warm:
                        ev[i][0].record(streams[i])
                    for _ in range(2 if warm else reps):
                        dev[i].copy_(host[i], non_blocking=True)
                        if both:
                            back[i].copy_(dev[i][: nbytes // 4], non_blocking=True)
                    if not warm:
                        ev[i][1].record(streams[i])
            for i in range(k):
                torch.cuda.synchronize(i)
        ms = max(a.elapsed_time(b) for a, b in ev) / reps
        out["h2d_plus_d2h" if both else "h2d"][str(k)] = {"ms_per_205MB_step": round(ms, 3), "aggregate_h2d_GBps": round(k * nbytes / ms / 1e6, 1), "per_device_GBps": round(nbytes / ms / 1e6, 1)}
print(json.dumps(out))
