#!/usr/bin/env python
"""Time the dense SparsePseudoInputGaussianProcess path at the size of BASELINE.json configs[4] ("SPGP occupancy map with
M = 2048 pseudo-inputs"; settings of the reference's config/spgp_occupancy_map_2d.yaml: Matern32 scale 0.18, noise 1e-4,
2000 samples per update) with the CPU port beside it on a shorter run.

    python tools/bench_spgp.py [--m 2048] [--updates 50] [--samples 2000] [--grid 100] [--dtype f64] [--cpu-updates 2]

Prints one JSON line: ms per Update() (K_MN, beta = L_KM^-1 K_MN, Q_M += beta beta^T / var, alpha accumulation), ms of the first
Test() (L_QM factorisation + predict of grid^2 points) and of a second Test() (predict only), useful TFLOP/s of the updates
(3 M^2 N per update), and the same for the OpenMP port.  Times are host wall-clock around the synchronous C-ABI calls."""
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
    ap.add_argument("--m", type=int, default=2048)
    ap.add_argument("--updates", type=int, default=50)
    ap.add_argument("--samples", type=int, default=2000)
    ap.add_argument("--grid", type=int, default=100)
    ap.add_argument("--dtype", default="f64", choices=["f32", "f64"])
    ap.add_argument("--cpu-updates", type=int, default=2, help="updates timed on the CPU port (0 = skip)")
    args = ap.parse_args()

    import erl_gaussian_process_b200 as gp

    dt = np.float64 if args.dtype == "f64" else np.float32
    side = int(round(np.sqrt(args.m / 2)))
    gx, gy = np.linspace(-3, 3, 2 * side), np.linspace(-3, 3, side)
    z = np.array([[a, b] for a in gx for b in gy])[: args.m]
    rng = np.random.default_rng(7)
    xs = [rng.uniform(-3, 3, (args.samples, 2)) for _ in range(args.updates)]
    ys = [np.tanh(x[:, 0] * x[:, 1]) for x in xs]
    var = np.full(args.samples, 1e-4)
    gt = np.linspace(-3, 3, args.grid)
    xt = np.array([[a, b] for a in gt for b in gt])

    ctx = gp.Context(0)
    g = gp.SparsePseudoInputGaussianProcess("matern32", 0.18, z, dt, ctx)
    g.update(xs[0], ys[0], var)  # warm-up (allocations)
    ctx.synchronize()
    t0 = time.perf_counter()
    for x, y in zip(xs, ys):
        g.update(x, y, var)
    ctx.synchronize()
    upd_ms = 1e3 * (time.perf_counter() - t0) / args.updates
    t0 = time.perf_counter()
    mean, variance = g.test(xt)
    test1_ms = 1e3 * (time.perf_counter() - t0)
    t0 = time.perf_counter()
    mean, variance = g.test(xt)
    test2_ms = 1e3 * (time.perf_counter() - t0)
    m = z.shape[0]
    out = {"m": m, "updates": args.updates, "samples": args.samples, "test_points": len(xt), "dtype": args.dtype, "gpu_update_ms": upd_ms,
           "gpu_update_tflops": 3.0 * m * m * args.samples / (upd_ms * 1e-3) / 1e12, "gpu_first_test_ms": test1_ms, "gpu_test_ms": test2_ms,
           "finite": bool(np.isfinite(mean).all() and np.isfinite(variance).all())}
    if args.cpu_updates > 0:
        import oracle

        o = oracle.Spgp(oracle.MATERN32, 0.18, z, dt)
        t0 = time.perf_counter()
        for x, y in zip(xs[: args.cpu_updates], ys[: args.cpu_updates]):
            o.update(x, y, var)
        out["cpu_update_ms"] = 1e3 * (time.perf_counter() - t0) / args.cpu_updates
        t0 = time.perf_counter()
        o.test(xt)
        out["cpu_first_test_ms"] = 1e3 * (time.perf_counter() - t0)
        out["cpu_cores"] = len(os.sched_getaffinity(0))
        # parity on the same (shorter) update sequence
        g2 = gp.SparsePseudoInputGaussianProcess("matern32", 0.18, z, dt, ctx)
        for x, y in zip(xs[: args.cpu_updates], ys[: args.cpu_updates]):
            g2.update(x, y, var)
        m2, v2 = g2.test(xt)
        mr, vr = o.test(xt)
        out["err_mean"] = float(np.abs(m2 - mr).max() / np.abs(mr).max())
        out["err_var"] = float(np.abs(v2 - vr).max())
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
