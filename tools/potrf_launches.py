#!/usr/bin/env python
"""One dense train (Gram + blocked Cholesky + alpha) of size n: used under `ncu --metrics gpu__time_duration.sum` to
list the per-kernel times of the factorisation.  usage: tools/potrf_launches.py <n> [f64|f32]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import erl_gaussian_process_b200 as gp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dt = np.float32 if len(sys.argv) > 2 and sys.argv[2] == "f32" else np.float64
rng = np.random.default_rng(1)
x = rng.uniform(-1, 1, (n, 2)).astype(dt)
y = (2 * np.sin(10 * x[:, 0]) * np.cos(10 * x[:, 1])).astype(dt)
var = np.full(n, 1e-3, dtype=dt)
ctx = gp.Context(0)
g = gp.VanillaGaussianProcess(gp.VanillaGaussianProcess.Setting("matern32", 0.1, -1), dt, ctx)
for _ in range(2):
    assert g.train(x, y, var)
ctx.synchronize()
print("ok", g.info, ctx.kernel_launches)
