"""Shared helpers for the parity tests: synthetic workloads and the parity metric.

Parity metric: norm-wise, because pointwise-relative error is not a usable
gate (two correct f64 implementations already differ by 1e-9 where the mean crosses zero):
    err_mean = max|d| / max(|ref|_inf, eps)        err_var = max|d| / 1.0   (prior variance is the literal 1)
Tolerances are the north_star's: 1e-10 in double, 1e-4 in float.
"""
import numpy as np

TOL = {np.dtype(np.float32): 1e-4, np.dtype(np.float64): 1e-10}


def err_mean(got, ref):
    ref = np.asarray(ref, dtype=np.float64)
    got = np.asarray(got, dtype=np.float64)
    return float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-300))


def err_var(got, ref):
    return float(np.abs(np.asarray(got, dtype=np.float64) - np.asarray(ref, dtype=np.float64)).max())


def make_batch(rng, num_gps, max_n, x_dim, dtype, n_lo=None, n_hi=None, q_lo=0, q_hi=64, fixed_q=None):
    """Random independent GPs: x ~ U[0,1]^d, y = smooth sum of sines, var = 0.01."""
    n_hi = max_n if n_hi is None else n_hi
    n_lo = n_hi if n_lo is None else n_lo
    n_train = rng.integers(n_lo, n_hi + 1, num_gps).astype(np.int32)
    x = rng.uniform(0, 1, (num_gps, max_n, x_dim))
    w = rng.uniform(1, 4, (num_gps, 1, x_dim))
    y = 0.5 * np.sin(w * x * 3.0).sum(axis=2)
    var = np.full((num_gps, max_n), 0.01)
    if fixed_q is not None:
        nq = np.full(num_gps, fixed_q, dtype=np.int64)
    else:
        nq = rng.integers(q_lo, q_hi + 1, num_gps).astype(np.int64)
    q_offsets = np.concatenate([[0], np.cumsum(nq)]).astype(np.int64)
    q_x = rng.uniform(0, 1, (int(q_offsets[-1]), x_dim))
    return n_train, x.astype(dtype), y.astype(dtype), var.astype(dtype), q_offsets, q_x.astype(dtype)
