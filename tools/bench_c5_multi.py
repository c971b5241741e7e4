"""C5 on the GPUs of one box from ONE process: VanillaGp<double> N = 16384 is factorised once on device 0, the trained
state (x_train, L, alpha: 2 GiB) is fanned out to the other devices with erl_gp_vanilla_replicate (cudaMemcpyPeerAsync over NVLink)
and the T test points are predicted with erl_gp_vanilla_test_multi over k = 1, 2, 4, 8 replicas.  Wall-clock around the synchronous
C-ABI calls (host buffers in, host buffers out).  Prints one JSON line."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

import erl_gaussian_process_b200 as gp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
t = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
ndev = gp._capi.device_count()
rng = np.random.default_rng(1)
x = rng.uniform(-1, 1, (n, 2))
y = 2 * np.sin(10 * x[:, 0]) * np.cos(10 * x[:, 1])
var = np.full(n, 1e-3)
xt = np.random.default_rng(2).uniform(-1, 1, (t, 2))
setting = gp.VanillaGaussianProcess.Setting("matern32", 0.1, max_num_samples=n)
g0 = gp.VanillaGaussianProcess(setting, np.float64, gp.Context(0))
g0.train(x, y, var)  # warm-up (allocations)
t0 = time.perf_counter()
g0.train(x, y, var)
train_ms = (time.perf_counter() - t0) * 1e3
replicas, rep_ms = [g0], []
for dev in range(1, ndev):
    r = gp.VanillaGaussianProcess(setting, np.float64, gp.Context(dev))
    g0.replicate_to(r)  # warm-up (allocations, peer access)
    t0 = time.perf_counter()
    g0.replicate_to(r)
    rep_ms.append((time.perf_counter() - t0) * 1e3)
    replicas.append(r)
out = {"workload": f"VanillaGp<double> N={n} Matern32(0.1), T={t} test points, one process", "devices": ndev, "train_ms_device0": round(train_ms, 2),
       "replicate_ms_per_device": [round(v, 2) for v in rep_ms], "state_bytes": n * n * 8 + n * 8 * 3, "predict": {}}
ref = None
for k in [v for v in (1, 2, 4, 8) if v <= ndev]:
    gp.VanillaGaussianProcess.test_multi(replicas[:k], xt[: max(4096 * k, 1)])  # warm-up
    t0 = time.perf_counter()
    mean, variance = gp.VanillaGaussianProcess.test_multi(replicas[:k], xt)
    dt = time.perf_counter() - t0
    if ref is None:
        ref = (mean, variance)
    same = bool(np.array_equal(mean, ref[0]) and np.abs(variance - ref[1]).max() < 1e-13)
    out["predict"][str(k)] = {"seconds": round(dt, 4), "test_points_per_s": round(t / dt, 1), "tflops_fp64": round(t * (n * n + 2.0 * n) / dt / 1e12, 2), "same_as_1_gpu": same}
print(json.dumps(out))
