"""ORACLE — TEST INFRASTRUCTURE ONLY.

Second, independent restatement of the reference hot path in numpy/scipy (LAPACK potrf/trtrs).
It exists so the C++ oracle (erl_gp_oracle.hpp) is not its own only witness: tests compare the
two with each other and with the reference's known-answer values
(test/gtest/test_vanilla_gp.cpp:103,214,366-367; test_sparse_pseudo_input_gp.cpp:109).

Nothing under erl_gaussian_process_b200/ imports this module.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import cholesky, solve_triangular

OU, MATERN32, RBF = 0, 1, 2


def kernel_from_r2(kernel: int, scale: float, r2: np.ndarray) -> np.ndarray:
    """erl_covariance v0.2.0 kernels, SURVEY.md Appendix A."""
    dt = r2.dtype.type
    if kernel == RBF:
        return np.exp(-r2 / (dt(2) * dt(scale) * dt(scale)))
    r = np.sqrt(r2)
    if kernel == MATERN32:
        ar = (np.sqrt(dt(3)) / dt(scale)) * r
        return (dt(1) + ar) * np.exp(-ar)
    return np.exp(-r / dt(scale))


def sqdist(x1: np.ndarray, x2: np.ndarray) -> np.ndarray:
    """x1: (n1, d), x2: (n2, d) -> (n1, n2) squared distances by explicit differences."""
    d = x1[:, None, :] - x2[None, :, :]
    return np.einsum("ijk,ijk->ij", d, d)


def ktrain(kernel: int, scale: float, x: np.ndarray, var: np.ndarray) -> np.ndarray:
    k = kernel_from_r2(kernel, scale, sqdist(x, x))
    k[np.diag_indices_from(k)] = x.dtype.type(1) + var
    return k


def ktest(kernel: int, scale: float, x: np.ndarray, xt: np.ndarray) -> np.ndarray:
    return kernel_from_r2(kernel, scale, sqdist(x, xt))


def vanilla_train(kernel, scale, x, y, var):
    """x: (n, d), y: (n,) or (n, ydim), var: (n,).  Returns (L, alpha).  src/vanilla_gp.cpp:476-505."""
    k = ktrain(kernel, scale, x, var)
    l = cholesky(k, lower=True)
    z = solve_triangular(l, y, lower=True)
    alpha = solve_triangular(l.T, z, lower=False)
    return l, alpha


def vanilla_test(kernel, scale, x, l, alpha, xt):
    """Returns (mean, var).  src/vanilla_gp.cpp:61-150: var = 1 - ||L^-1 k*||^2."""
    kt = ktest(kernel, scale, x, xt)
    mean = kt.T @ alpha
    v = solve_triangular(l, kt, lower=True)
    var = x.dtype.type(1) - np.einsum("ij,ij->j", v, v)
    return mean, var


def make_partitions(coords, group_size, overlap_size, margin, symmetric=True):
    """src/lidar_gp_2d.cpp:238-300 / src/range_sensor_gp_3d.cpp:199-259."""
    n = len(coords)
    step = group_size - overlap_size
    g = max(1, n // step) + 1
    gs2 = (n - (g - 2) * step) // 2
    ho = overlap_size // 2
    parts = []
    if symmetric:
        parts.append((0, gs2 + ho, coords[margin], coords[gs2]))
        for i in range(g - 2):
            il = i * step + gs2 - ho
            ir = il + group_size
            parts.append((il, ir, coords[il + ho], coords[ir - ho]))
        parts.append((n - gs2 - ho, n, coords[n - 1 - gs2], coords[n - 1 - margin]))
        return parts
    for i in range(g - 2):
        il = i * step
        ir = il + group_size
        parts.append((il, ir, coords[il], coords[ir - ho]))
    il = (g - 2) * step
    ir = il + (n - il + overlap_size) // 2
    parts.append((il, ir, coords[il], coords[ir - ho]))
    il = il + (n - il - overlap_size) // 2
    parts.append((il, n, coords[il], coords[n - 1]))
    return parts


def make_hit_ray_partitions(angles, mask_hit, group_size, overlap_size):
    """src/lidar_gp_2d.cpp:302-348, written literally (Python raises IndexError where the reference reads out of
    bounds) when clamp=False semantics are wanted use `strict_hit_ray_partitions`; this one clamps like the C++ oracle."""
    return _hit_ray_partitions(angles, mask_hit, group_size, overlap_size, clamp=True)


def strict_hit_ray_partitions(angles, mask_hit, group_size, overlap_size):
    """The same formulas without any clamp: raises IndexError exactly where the reference's reads are undefined."""
    return _hit_ray_partitions(angles, mask_hit, group_size, overlap_size, clamp=False)


def _hit_ray_partitions(angles, mask_hit, group_size, overlap_size, clamp):
    hit = np.flatnonzero(np.asarray(mask_hit))
    n, num_rays = len(hit), len(angles)
    if n == 0:
        return None  # "No hit rays are stored": the table is left as it was (:306-309)

    def h(i):
        if clamp:
            i = min(max(i, 0), n - 1)
        elif not 0 <= i < n:
            raise IndexError(f"hit_ray_indices[{i}] with {n} hit rays")
        return int(hit[i])

    def a(i):
        if clamp:
            i = min(max(i, 0), num_rays - 1)
        elif not 0 <= i < num_rays:
            raise IndexError(f"angles[{i}] with {num_rays} rays")
        return angles[i]

    def tdiv(p, q):  # C++ integer division truncates toward zero
        return int(p / q) if p < 0 else p // q

    step = group_size - overlap_size
    g = max(1, n // step) + 1
    parts = []
    for i in range(g - 2):
        il, ir = h(i * step), h(i * step + group_size)
        parts.append((il, ir, a(il), a(ir)))
    il = (g - 2) * step
    ir = il + tdiv(n - il + overlap_size, 2)
    il, ir = h(il), h(ir)
    parts.append((il, ir, a(il), a(ir)))
    il = il + tdiv(n - il - overlap_size, 2)  # :341 — an original ray index used as a hit-ray position
    il = h(il)
    ir = h(n - 1) + 1
    parts.append((il, ir, a(il), a(ir)))
    return parts


def spgp_fit_predict(kernel, scale, z, x, y, var, xt):
    """Dense SPGP, one update then predict.  src/sparse_pseudo_input_gp.cpp:313-356, 751-791, 43-113, 280-310."""
    one = z.dtype.type(1)
    k_m = kernel_from_r2(kernel, scale, sqdist(z, z))
    l_km = cholesky(k_m, lower=True)
    k_mn = kernel_from_r2(kernel, scale, sqdist(z, x))
    beta = solve_triangular(l_km, k_mn, lower=True)
    lam = one - np.einsum("ij,ij->j", beta, beta)
    k_s = k_mn / (lam + var)[None, :]
    q_m = k_m + k_s @ k_mn.T
    alpha = k_s @ y
    l_qm = cholesky(q_m, lower=True)
    a = solve_triangular(l_qm.T, solve_triangular(l_qm, alpha, lower=True), lower=False)
    k_t = kernel_from_r2(kernel, scale, sqdist(z, xt))
    mean = k_t.T @ a
    b = solve_triangular(l_km, k_t, lower=True)
    g = solve_triangular(l_qm, k_t, lower=True)
    variance = one - np.einsum("ij,ij->j", b, b) + np.einsum("ij,ij->j", g, g)
    return mean, variance
