#!/usr/bin/env python
"""Randomised stress of the batched kernels (compute-sanitizer is not available on the pool): many ragged GPs against the CPU
port, plus run-to-run bitwise determinism of every output (a shared-memory race would show up as a difference).

    python tools/stress_batch.py [--gps 3000] [--rounds 3]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import make_batch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gps", type=int, default=3000)
    ap.add_argument("--rounds", type=int, default=3)
    args = ap.parse_args()
    import erl_gaussian_process_b200 as gp
    import oracle

    worst = {}
    for dtype, tol in ((np.float32, 1e-4), (np.float64, 1e-10)):
        for max_n, x_dim, kernel, okern, scale in ((128, 3, "matern32", oracle.MATERN32, 0.3), (64, 1, "ou", oracle.OU, 0.05), (192, 2, "matern32", oracle.MATERN32, 0.3),
                                                   (256 if dtype == np.float32 else 160, 2, "rbf", oracle.RBF, 0.25)):
            rng = np.random.default_rng(max_n + x_dim)
            batch = make_batch(rng, args.gps if max_n <= 128 else args.gps // 4, max_n, x_dim, dtype, n_lo=0, n_hi=max_n, q_lo=0, q_hi=200)
            ref = oracle.batched_train_predict(okern, scale, *batch)
            b = gp.BatchGp(len(batch[0]), max_n, x_dim, kernel, scale, dtype)
            first = None
            for r in range(args.rounds):
                out = b.train_predict(*batch)
                keys = ("mean", "var", "valid", "info", "alpha")
                if first is None:
                    first = {k: np.array(out[k], copy=True) for k in keys}
                else:
                    for k in keys:
                        assert np.array_equal(first[k], out[k], equal_nan=True), f"{dtype.__name__} n<={max_n} {kernel}: output '{k}' differs between runs"
            ok = out["valid"].astype(bool)
            assert np.array_equal(ok, np.asarray(ref["valid"]).astype(bool)) if "valid" in ref else True
            trained = np.asarray(ref["info"]) == 0
            assert np.array_equal(np.asarray(out["info"]) == 0, trained)
            em = np.abs(out["mean"][ok] - ref["mean"][ok]).max() / np.abs(ref["mean"][ok]).max()
            ev = np.abs(out["var"][ok] - ref["var"][ok]).max()
            worst[(dtype.__name__, max_n, kernel)] = (float(em), float(ev))
            assert em < tol and ev < tol, (dtype.__name__, max_n, kernel, em, ev)
    for k, v in worst.items():
        print(k, "mean err %.2e  var err %.2e" % v)
    print("stress ok")


if __name__ == "__main__":
    main()
