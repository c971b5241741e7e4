"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes front-end of oracle/_build/liberl_gp_oracle.so (the C++17/OpenMP restatement of the
reference hot path, oracle/erl_gp_oracle.hpp).  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liberl_gp_oracle.so")

OU, MATERN32, RBF = 0, 1, 2
KERNELS = {"ou": OU, "matern32": MATERN32, "rbf": RBF}


def build(force: bool = False) -> str:
    if force or not os.path.exists(_SO):
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _sfx(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32", C.c_float
    if dtype == np.float64:
        return "f64", C.c_double
    raise TypeError(dtype)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _fn(name, dtype, restype=C.c_int):
    sfx, _ = _sfx(dtype)
    f = getattr(lib(), f"{name}_{sfx}")
    f.restype = restype
    return f


def num_threads() -> int:
    return lib().oracle_num_threads()


def set_num_threads(n: int) -> None:
    lib().oracle_set_num_threads(int(n))


def gram_train(kernel, scale, x, var):
    """x: (n, d) C-contiguous (== d x n col-major); returns K as (n, n) (symmetric)."""
    x = np.ascontiguousarray(x)
    n, d = x.shape
    _, ct = _sfx(x.dtype)
    k = np.zeros((n, n), dtype=x.dtype)
    _fn("oracle_gram_train", x.dtype)(C.c_int(kernel), ct(scale), C.c_long(d), _p(x), C.c_long(d), _p(np.ascontiguousarray(var)), C.c_long(n), _p(k), C.c_long(n))
    return k


def gram_test(kernel, scale, x1, x2):
    """Returns Ktest as an (n1, n2) array (element [i, j] = k(x1_i, x2_j))."""
    x1 = np.ascontiguousarray(x1)
    x2 = np.ascontiguousarray(x2)
    n1, d = x1.shape
    n2 = x2.shape[0]
    _, ct = _sfx(x1.dtype)
    kt = np.zeros((n2, n1), dtype=x1.dtype)  # col-major n1 x n2
    _fn("oracle_gram_test", x1.dtype)(C.c_int(kernel), ct(scale), C.c_long(d), _p(x1), C.c_long(d), C.c_long(n1), _p(x2), C.c_long(d), C.c_long(n2), _p(kt), C.c_long(n1))
    return kt.T


def llt(k):
    """k: (n, n) symmetric.  Returns (L as (n, n) lower, info)."""
    k = np.ascontiguousarray(k)
    n = k.shape[0]
    l = np.zeros((n, n), dtype=k.dtype)
    info = _fn("oracle_llt", k.dtype)(_p(k), C.c_long(n), C.c_long(n), _p(l), C.c_long(n))
    return l.T.copy(), info  # col-major lower -> numpy [i, j]


class VanillaGp:
    """oracle VanillaGaussianProcess (src/vanilla_gp.cpp)."""

    def __init__(self, kernel, scale, dtype=np.float64, max_num_samples=-1):
        self.dtype = np.dtype(dtype)
        _, self.ct = _sfx(dtype)
        self.h = C.c_void_p(_fn("oracle_vanilla_create", dtype, C.c_void_p)(C.c_int(kernel), self.ct(scale), C.c_long(max_num_samples)))
        self.n = 0
        self.y_dim = 1

    def __del__(self):
        if getattr(self, "h", None):
            _fn("oracle_vanilla_destroy", self.dtype, None)(self.h)
            self.h = None

    def train(self, x, y, var):
        x = np.ascontiguousarray(x, dtype=self.dtype)
        y = np.asarray(y, dtype=self.dtype)
        if y.ndim == 1:
            y = y[:, None]
        yf = np.asfortranarray(y)
        var = np.ascontiguousarray(var, dtype=self.dtype)
        n, d = x.shape
        self.n, self.y_dim = n, y.shape[1]
        return _fn("oracle_vanilla_train", self.dtype)(self.h, C.c_long(n), C.c_long(d), C.c_long(self.y_dim), _p(x), C.c_long(d), _p(yf), C.c_long(n), _p(var))

    def get(self):
        n = self.n
        k = np.zeros((n, n), dtype=self.dtype)
        l = np.zeros((n, n), dtype=self.dtype)
        a = np.zeros((self.y_dim, n), dtype=self.dtype)
        _fn("oracle_vanilla_get", self.dtype)(self.h, _p(k), C.c_long(n), _p(l), C.c_long(n), _p(a), C.c_long(n))
        return k.T.copy(), l.T.copy(), a.T.copy()

    def test(self, xt, parallel=True, want_var=True):
        xt = np.ascontiguousarray(xt, dtype=self.dtype)
        t, d = xt.shape
        mean = np.zeros((self.y_dim, t), dtype=self.dtype)
        var = np.zeros(t, dtype=self.dtype) if want_var else None
        rc = _fn("oracle_vanilla_test", self.dtype)(self.h, _p(xt), C.c_long(d), C.c_long(t), _p(mean), _p(var), C.c_int(int(parallel)))
        assert rc == 0
        return (mean[0] if self.y_dim == 1 else mean.T.copy()), var


def make_partitions(coords, group_size, overlap_size, margin, symmetric=True):
    coords = np.ascontiguousarray(coords)
    cap = len(coords) + 2
    il = np.zeros(cap, dtype=np.int64)
    ir = np.zeros(cap, dtype=np.int64)
    cl = np.zeros(cap, dtype=coords.dtype)
    cr = np.zeros(cap, dtype=coords.dtype)
    n = _fn("oracle_make_partitions", coords.dtype, C.c_long)(_p(coords), C.c_long(1), C.c_long(len(coords)), C.c_long(group_size), C.c_long(overlap_size), C.c_long(margin), C.c_int(int(symmetric)), C.c_long(cap), _p(il), _p(ir), _p(cl), _p(cr))
    return [(int(il[i]), int(ir[i]), cl[i], cr[i]) for i in range(n)]


class LidarGp2D:
    """oracle LidarGaussianProcess2D (src/lidar_gp_2d.cpp); the LidarFrame2D outputs are inputs."""

    def __init__(self, angles, kernel=OU, scale=1.0, group_size=26, overlap_size=6, margin=1, symmetric=True,
                 sensor_range_var=0.01, discontinuity_var=10.0, discontinuity_detection=False, mapping_type=2,
                 mapping_scale=1.0, max_valid_range_var=0.1, occ_test_temperature=30.0, dtype=np.float64, partition_on_hit_rays=False):
        self.dtype = np.dtype(dtype)
        _, ct = _sfx(dtype)
        self.ct = ct
        self.angles = np.ascontiguousarray(angles, dtype=self.dtype)
        self.group_size = group_size
        self.h = C.c_void_p(_fn("oracle_lidar_create", dtype, C.c_void_p)(
            C.c_int(kernel), ct(scale), C.c_long(group_size), C.c_long(overlap_size), C.c_long(margin), C.c_int(int(symmetric)),
            ct(sensor_range_var), ct(discontinuity_var), C.c_int(int(discontinuity_detection)), C.c_int(mapping_type), ct(mapping_scale),
            ct(max_valid_range_var), ct(occ_test_temperature), _p(self.angles), C.c_long(len(self.angles))))
        if partition_on_hit_rays:  # src/lidar_gp_2d.cpp:182, 302-348, 364
            _fn("oracle_lidar_set_partition_on_hit_rays", dtype, None)(self.h, C.c_int(1))

    @property
    def angle_partitions(self):
        n = self.num_partitions
        il, ir = np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.int64)
        cl, cr = np.zeros(n, dtype=self.dtype), np.zeros(n, dtype=self.dtype)
        _fn("oracle_lidar_partitions", self.dtype, C.c_long)(self.h, _p(il), _p(ir), _p(cl), _p(cr))
        return [(int(il[i]), int(ir[i]), cl[i], cr[i]) for i in range(n)]

    def __del__(self):
        if getattr(self, "h", None):
            _fn("oracle_lidar_destroy", self.dtype, None)(self.h)
            self.h = None

    @property
    def num_partitions(self):
        return _fn("oracle_lidar_num_partitions", self.dtype, C.c_long)(self.h)

    def train(self, ranges, mask_hit, mask_con=None, rotation=None, frame_valid=True):
        ranges = np.ascontiguousarray(ranges, dtype=self.dtype)
        mask_hit = np.ascontiguousarray(mask_hit, dtype=np.uint8)
        mask_con = np.ones_like(mask_hit) if mask_con is None else np.ascontiguousarray(mask_con, dtype=np.uint8)
        rot = np.eye(2, dtype=self.dtype) if rotation is None else np.asarray(rotation, dtype=self.dtype)
        rot = np.asfortranarray(rot)
        return _fn("oracle_lidar_train", self.dtype)(self.h, _p(rot), _p(ranges), _p(mask_hit), _p(mask_con), C.c_int(int(frame_valid))) == 0

    def get_gp(self, p):
        gs = self.group_size
        trained = C.c_int(0)
        n = C.c_long(0)
        l = np.zeros((gs, gs), dtype=self.dtype)
        a = np.zeros(gs, dtype=self.dtype)
        _fn("oracle_lidar_get_gp", self.dtype)(self.h, C.c_long(p), C.byref(trained), C.byref(n), _p(l), C.c_long(gs), _p(a))
        return bool(trained.value), int(n.value), l.T[: n.value, : n.value].copy(), a[: n.value].copy()

    def test(self, angles, angles_are_local=True, un_map=True):
        angles = np.ascontiguousarray(angles, dtype=self.dtype)
        t = len(angles)
        mean = np.full(t, np.nan, dtype=self.dtype)
        var = np.full(t, np.nan, dtype=self.dtype)
        valid = np.zeros(t, dtype=np.uint8)
        rc = _fn("oracle_lidar_test", self.dtype)(self.h, _p(angles), C.c_long(t), C.c_int(int(angles_are_local)), C.c_int(int(un_map)), _p(mean), _p(var), _p(valid))
        assert rc == 0
        return mean, var, valid.astype(bool)

    def compute_occ(self, px, py):
        d, r, o = self.ct(0), self.ct(0), self.ct(0)
        ok = _fn("oracle_lidar_compute_occ", self.dtype)(self.h, self.ct(px), self.ct(py), C.byref(d), C.byref(r), C.byref(o))
        return bool(ok), d.value, r.value, o.value


class RangeSensorGp3D:
    """oracle RangeSensorGaussianProcess3D (src/range_sensor_gp_3d.cpp)."""

    def __init__(self, frame_coords, kernel=OU, scale=1.0, row_group_size=24, row_overlap_size=6, row_margin=0,
                 col_group_size=8, col_overlap_size=2, col_margin=0, min_num_samples_per_group=32, sensor_range_var=0.01,
                 mapping_type=2, mapping_scale=1.0, dtype=np.float32):
        """frame_coords: (rows, cols, 2) numpy array."""
        self.dtype = np.dtype(dtype)
        _, ct = _sfx(dtype)
        fc = np.asarray(frame_coords, dtype=self.dtype)
        self.rows, self.cols = fc.shape[:2]
        # Eigen col-major matrix of Vector2: index (r + c*rows)*2 + k  == array [c, r, k]
        self.fc = np.ascontiguousarray(fc.transpose(1, 0, 2))
        self.max_n = row_group_size * col_group_size
        h = _fn("oracle_range3d_create", dtype, C.c_void_p)(
            C.c_int(kernel), ct(scale), C.c_long(row_group_size), C.c_long(row_overlap_size), C.c_long(row_margin), C.c_long(col_group_size),
            C.c_long(col_overlap_size), C.c_long(col_margin), C.c_long(min_num_samples_per_group), ct(sensor_range_var), C.c_int(mapping_type),
            ct(mapping_scale), _p(self.fc), C.c_long(self.rows), C.c_long(self.cols))
        if not h:
            raise ValueError("overlap sizes must be even (src/range_sensor_gp_3d.cpp:190-197)")
        self.h = C.c_void_p(h)

    def __del__(self):
        if getattr(self, "h", None):
            _fn("oracle_range3d_destroy", self.dtype, None)(self.h)
            self.h = None

    @property
    def grid(self):
        a, b = C.c_long(0), C.c_long(0)
        _fn("oracle_range3d_grid", self.dtype, None)(self.h, C.byref(a), C.byref(b))
        return int(a.value), int(b.value)

    def train(self, ranges, mask_hit, frame_valid=True):
        """ranges, mask_hit: (rows, cols)."""
        r = np.asfortranarray(np.asarray(ranges, dtype=self.dtype))
        m = np.asfortranarray(np.asarray(mask_hit, dtype=np.uint8))
        return _fn("oracle_range3d_train", self.dtype)(self.h, _p(r), _p(m), C.c_int(int(frame_valid))) == 0

    def get_gp(self, g):
        mn = self.max_n
        trained = C.c_int(0)
        n = C.c_long(0)
        l = np.zeros((mn, mn), dtype=self.dtype)
        a = np.zeros(mn, dtype=self.dtype)
        _fn("oracle_range3d_get_gp", self.dtype)(self.h, C.c_long(g), C.byref(trained), C.byref(n), _p(l), C.c_long(mn), _p(a))
        return bool(trained.value), int(n.value), l.T[: n.value, : n.value].copy(), a[: n.value].copy()

    def test(self, coords, coords_ok=None, un_map=True):
        """coords: (T, 2) frame coordinates."""
        coords = np.ascontiguousarray(coords, dtype=self.dtype)
        t = coords.shape[0]
        ok = None if coords_ok is None else np.ascontiguousarray(coords_ok, dtype=np.uint8)
        mean = np.full(t, np.nan, dtype=self.dtype)
        var = np.full(t, np.nan, dtype=self.dtype)
        valid = np.zeros(t, dtype=np.uint8)
        rc = _fn("oracle_range3d_test", self.dtype)(self.h, _p(coords), _p(ok), C.c_long(t), C.c_int(int(un_map)), _p(mean), _p(var), _p(valid))
        assert rc == 0
        return mean, var, valid.astype(bool)


def batched_train_predict(kernel, scale, n_train, x, y, var, q_offsets=None, q_x=None, want_l=True):
    """x: (B, max_n, d); y, var: (B, max_n); n_train: (B,) int32; q_offsets: (B+1,) int64; q_x: (T, d).
    Returns dict(L (B, max_n, max_n) with [g, r, c] indexing, alpha, info, mean, var)."""
    x = np.ascontiguousarray(x)
    dt = x.dtype
    _, ct = _sfx(dt)
    b, max_n, d = x.shape
    y = np.ascontiguousarray(y, dtype=dt)
    var = np.ascontiguousarray(var, dtype=dt)
    n_train = np.ascontiguousarray(n_train, dtype=np.int32)
    l = np.zeros((b, max_n, max_n), dtype=dt) if want_l else None
    alpha = np.zeros((b, max_n), dtype=dt)
    info = np.zeros(b, dtype=np.int32)
    mean = varo = None
    if q_offsets is not None:
        q_offsets = np.ascontiguousarray(q_offsets, dtype=np.int64)
        q_x = np.ascontiguousarray(q_x, dtype=dt)
        t = q_x.shape[0]
        mean = np.full(t, np.nan, dtype=dt)
        varo = np.full(t, np.nan, dtype=dt)
    _fn("oracle_batched_train_predict", dt)(C.c_int(kernel), ct(scale), C.c_long(d), C.c_long(b), C.c_long(max_n), _p(n_train), _p(x), _p(y), _p(var),
                                            _p(q_offsets), _p(q_x), _p(l), _p(alpha), _p(info), _p(mean), _p(varo))
    return dict(L=None if l is None else l.transpose(0, 2, 1), alpha=alpha, info=info, mean=mean, var=varo)


class Spgp:
    """oracle SparsePseudoInputGaussianProcess, dense mode (src/sparse_pseudo_input_gp.cpp)."""

    def __init__(self, kernel, scale, pseudo, dtype=np.float64):
        self.dtype = np.dtype(dtype)
        _, ct = _sfx(dtype)
        z = np.ascontiguousarray(pseudo, dtype=self.dtype)
        self.m, self.d = z.shape
        self.h = C.c_void_p(_fn("oracle_spgp_create", dtype, C.c_void_p)(C.c_int(kernel), ct(scale), C.c_long(self.d), C.c_long(self.m), _p(z)))

    def __del__(self):
        if getattr(self, "h", None):
            _fn("oracle_spgp_destroy", self.dtype, None)(self.h)
            self.h = None

    def update(self, x, y, var):
        x = np.ascontiguousarray(x, dtype=self.dtype)
        y = np.ascontiguousarray(y, dtype=self.dtype)
        var = np.ascontiguousarray(var, dtype=self.dtype)
        return _fn("oracle_spgp_update", self.dtype)(self.h, _p(x), _p(y), _p(var), C.c_long(x.shape[0])) == 0

    def test(self, xt):
        xt = np.ascontiguousarray(xt, dtype=self.dtype)
        t = xt.shape[0]
        mean = np.zeros(t, dtype=self.dtype)
        var = np.zeros(t, dtype=self.dtype)
        _fn("oracle_spgp_test", self.dtype)(self.h, _p(xt), C.c_long(t), _p(mean), _p(var))
        return mean, var

    def set_diagonal_qm(self, on=True):
        _fn("oracle_spgp_set_diagonal_qm", self.dtype, None)(self.h, C.c_int(int(on)))

    def get_qm_diagonal(self):
        q = np.zeros(self.m, dtype=self.dtype)
        _fn("oracle_spgp_get_qm_diagonal", self.dtype)(self.h, _p(q))
        return q

    def test_gradient(self, xt, raw_alpha=False):
        xt = np.ascontiguousarray(xt, dtype=self.dtype)
        t = xt.shape[0]
        grad = np.zeros((t, self.d), dtype=self.dtype)
        _fn("oracle_spgp_test_gradient", self.dtype)(self.h, _p(xt), C.c_long(t), _p(grad), C.c_int(int(raw_alpha)))
        return grad

    def get(self):
        m = self.m
        q = np.zeros((m, m), dtype=self.dtype)
        a = np.zeros(m, dtype=self.dtype)
        lk = np.zeros((m, m), dtype=self.dtype)
        lq = np.zeros((m, m), dtype=self.dtype)
        _fn("oracle_spgp_get", self.dtype)(self.h, _p(q), _p(a), _p(lk), _p(lq))
        return q.T.copy(), a, lk.T.copy(), lq.T.copy()


class NoisyInputGp:
    """oracle NoisyInputGaussianProcess (src/noisy_input_gp.cpp): GP with noisy inputs and gradient observations.
    x: (n, x_dim); y: (n, y_dim); grad: (n, y_dim, x_dim) = d y_d / d x_k; grad_flag: (n,) ints."""

    def __init__(self, kernel, scale, no_gradient_observation=False, dtype=np.float64):
        self.dtype = np.dtype(dtype)
        _, ct = _sfx(dtype)
        self.h = C.c_void_p(_fn("oracle_noisy_create", dtype, C.c_void_p)(C.c_int(kernel), ct(scale), C.c_int(int(no_gradient_observation))))
        self.m = 0

    def __del__(self):
        if getattr(self, "h", None):
            _fn("oracle_noisy_destroy", self.dtype, None)(self.h)
            self.h = None

    def train(self, x, y, grad, var_x, var_y, var_grad, grad_flag):
        x = np.ascontiguousarray(x, dtype=self.dtype)
        n, self.x_dim = x.shape
        y = np.asarray(y, dtype=self.dtype).reshape(n, -1)
        self.y_dim = y.shape[1]
        yf = np.asfortranarray(y)  # n x y_dim col-major
        g = None if grad is None else np.ascontiguousarray(np.asarray(grad, dtype=self.dtype).reshape(n, self.y_dim * self.x_dim))  # column i of the reference's grad matrix
        vx, vy = (np.ascontiguousarray(np.broadcast_to(v, (n,)), dtype=self.dtype) for v in (var_x, var_y))
        vg = None if var_grad is None else np.ascontiguousarray(np.broadcast_to(var_grad, (n,)), dtype=self.dtype)
        flag = np.ascontiguousarray(np.broadcast_to(grad_flag, (n,)), dtype=np.int64)
        self.m = _fn("oracle_noisy_train", self.dtype, C.c_long)(self.h, C.c_long(n), C.c_long(self.x_dim), C.c_long(self.y_dim), _p(x), _p(yf), _p(g), _p(vx), _p(vy), _p(vg), _p(flag))
        return self.m > 0

    def get(self):
        m = self.m
        k, l = np.zeros((m, m), dtype=self.dtype), np.zeros((m, m), dtype=self.dtype)
        a = np.zeros((self.y_dim, m), dtype=self.dtype)
        info = _fn("oracle_noisy_get", self.dtype)(self.h, _p(k), _p(l), _p(a))
        return info, k.T.copy(), l.T.copy(), a.T.copy()

    def test(self, xt, predict_gradient=True, variance=True):
        """-> mean (T, y_dim), gradient (T, y_dim, x_dim), var (T), grad_var (T, x_dim), cov (T, x_dim (x_dim + 1) / 2)"""
        xt = np.ascontiguousarray(xt, dtype=self.dtype)
        t, d = xt.shape
        mean = np.zeros((self.y_dim, t), dtype=self.dtype)
        grad = np.zeros((self.y_dim, t, d), dtype=self.dtype) if predict_gradient else None
        var = np.zeros(t, dtype=self.dtype) if variance else None
        gvar = np.zeros((t, d), dtype=self.dtype) if variance and predict_gradient else None
        cov = np.zeros((t, d * (d + 1) // 2), dtype=self.dtype) if variance and predict_gradient else None
        rc = _fn("oracle_noisy_test", self.dtype)(self.h, _p(xt), C.c_long(t), C.c_int(int(predict_gradient)), _p(mean), _p(grad), _p(var), _p(gvar), _p(cov))
        assert rc == 0
        return mean.T.copy(), (None if grad is None else grad.transpose(1, 0, 2).copy()), var, gvar, cov
