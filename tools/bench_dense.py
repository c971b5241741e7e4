#!/usr/bin/env python
"""Time the dense VanillaGaussianProcess path (BASELINE.json configs[0] "c1" and configs[4] "c5").

    python tools/bench_dense.py --n 16384 --t 131072 --dtype f64 [--reps 3]

Prints one JSON line: train ms (Gram + blocked Cholesky + alpha), useful TFLOP/s of the factorisation
(n^3/3), test ms for T points (mean + variance) and the useful TFLOP/s of the predict solve (T n^2).
Timed on the host around the synchronous C-ABI calls (host<->device copies of x / y / x* / mean / var are
inside: a few MB, negligible beside n^3/3 at these sizes)."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", "--size", dest="n", type=int, default=16384)  # --size / --tests: torchrun's own parser trips over --n
    ap.add_argument("--t", "--tests", dest="t", type=int, default=131072)
    ap.add_argument("--dtype", default="f64", choices=["f32", "f64"])
    ap.add_argument("--scale", type=float, default=0.1)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--check", action="store_true", help="compare a 512-point slice with numpy/LAPACK (small n only)")
    args = ap.parse_args()

    import erl_gaussian_process_b200 as gp

    dt = np.float64 if args.dtype == "f64" else np.float32
    rng = np.random.default_rng(1)
    x = rng.uniform(-1, 1, (args.n, 2)).astype(dt)
    y = (2 * np.sin(10 * x[:, 0]) * np.cos(10 * x[:, 1])).astype(dt)
    var = np.full(args.n, 1e-3, dtype=dt)
    xt = np.random.default_rng(2).uniform(-1, 1, (args.t, 2)).astype(dt)

    # several GPUs (torchrun, one process per GPU): every rank factors its own replica (the factorisation is "replicas only",
    # SURVEY.md 8e) and predicts a contiguous range of the test points; rank 0 reports the aggregate (max over ranks of the time)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist

        from erl_gaussian_process_b200 import sharding

        torch.cuda.set_device(local)
        dist.init_process_group("nccl")
        t_begin, t_end = sharding.shard_range(args.t, rank, world)
        xt = np.ascontiguousarray(xt[t_begin:t_end])
    ctx = gp.Context(local)
    g = gp.VanillaGaussianProcess(gp.VanillaGaussianProcess.Setting("matern32", args.scale, -1), dt, ctx)
    train_ms, test_ms = [], []
    for _ in range(args.reps + 1):
        ctx.synchronize()
        t0 = time.perf_counter()
        assert g.train(x, y, var)
        ctx.synchronize()
        train_ms.append(1e3 * (time.perf_counter() - t0))
        res = g.test(xt)
        t0 = time.perf_counter()
        res._run(True, True)
        ctx.synchronize()
        test_ms.append(1e3 * (time.perf_counter() - t0))
    train = min(train_ms[1:])
    test = min(test_ms[1:])
    if world > 1:
        tt = torch.tensor([train, test], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        train, test = float(tt[0]), float(tt[1])
        dist.barrier()
        dist.destroy_process_group()
        if rank != 0:
            return
    out = {"n": args.n, "t": args.t, "n_gpus": world, "dtype": args.dtype, "info": g.info, "train_ms": train, "potrf_tflops": args.n ** 3 / 3 / (train * 1e-3) / 1e12, "test_ms": test,
           "predict_tflops": args.t * (args.n ** 2 + 2 * args.n) / (test * 1e-3) / 1e12, "test_points_per_s": args.t / (test * 1e-3), "train_ms_all": train_ms, "test_ms_all": test_ms,
           "launches": ctx.kernel_launches}
    if args.check:
        import scipy.linalg as sl

        from oracle import oracle_np

        k = oracle_np.ktrain(oracle_np.MATERN32, args.scale, x.astype(np.float64), var.astype(np.float64))
        c = sl.cho_factor(k, lower=True)
        alpha = sl.cho_solve(c, y.astype(np.float64))
        kt = oracle_np.ktest(oracle_np.MATERN32, args.scale, x.astype(np.float64), xt[:512].astype(np.float64))
        mean = kt.T @ alpha
        v = sl.solve_triangular(c[0], kt, lower=True)
        variance = 1.0 - (v * v).sum(axis=0)
        out["err_mean"] = float(np.abs(res._mean[0][:512] - mean).max() / np.abs(mean).max())
        out["err_var"] = float(np.abs(res._var[:512] - variance).max())
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
