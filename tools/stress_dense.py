#!/usr/bin/env python
"""Dense path stress: the look-ahead Cholesky runs on three streams and the alpha solve spins on flags - train several
ragged sizes repeatedly, require bit-identical L / alpha from run to run and agreement with LAPACK.

    python tools/stress_dense.py [--rounds 4]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=4)
    args = ap.parse_args()
    import scipy.linalg as sl

    import erl_gaussian_process_b200 as gp
    from oracle import oracle_np

    for dt, tol in ((np.float64, 1e-10), (np.float32, 2e-3)):
        for n, ydim in ((513, 1), (1100, 3), (2500, 1), (5000, 2), (6150, 1)):
            rng = np.random.default_rng(n)
            x = rng.uniform(-1, 1, (n, 2)).astype(dt)
            y = np.stack([np.sin(3 * x).sum(axis=1) * (c + 1) for c in range(ydim)], axis=1).astype(dt)
            var = rng.uniform(0.005, 0.02, n).astype(dt)
            g = gp.VanillaGaussianProcess(gp.VanillaGaussianProcess.Setting("matern32", 0.3, -1), dt)
            first = None
            for r in range(args.rounds):
                assert g.train(x, y if ydim > 1 else y[:, 0], var) and g.info == 0
                _, l, a = g.get()
                if first is None:
                    first = (l.copy(), a.copy())
                else:
                    assert np.array_equal(first[0], l), f"n={n} {dt.__name__}: L differs between runs"
                    assert np.array_equal(first[1], a), f"n={n} {dt.__name__}: alpha differs between runs"
            k = oracle_np.ktrain(oracle_np.MATERN32, 0.3, x.astype(np.float64), var.astype(np.float64))
            c = sl.cho_factor(k, lower=True)
            a_ref = sl.cho_solve(c, y.astype(np.float64))
            l_ref = np.tril(c[0])
            el = np.abs(first[0] - l_ref).max() / np.abs(l_ref).max()
            ea = np.abs(first[1].reshape(a_ref.shape) - a_ref).max() / np.abs(a_ref).max()
            print(f"{dt.__name__} n={n} y_dim={ydim}: L err {el:.2e}  alpha err {ea:.2e}")
            assert el < tol and ea < tol * 1e3, (n, el, ea)
    print("stress ok")


if __name__ == "__main__":
    main()
