"""Host-side mirror of the reference's operator interface over the C ABI.

Class and method names follow the reference (VanillaGaussianProcess.train/test,
LidarGaussianProcess2D.train/test, RangeSensorGaussianProcess3D.train/test — see
python/binding/bind_vanilla_gp.cpp:79-101, bind_lidar_gp_2d.cpp:77-95,
bind_range_sensor_gp_3d.cpp:80-112) so the parity tests read like the reference's own tests.
All numerics run in liberl_gp_b200.so; numpy arrays are only the host buffers the C ABI reads
and writes.  Array conventions: points are rows here (``x[i]`` = sample i, C-contiguous
``(n, x_dim)``), which is byte-identical to the reference's column-major ``x_dim x n``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import ErlGpError, check, load


def _sfx(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32", C.c_float
    if dtype == np.float64:
        return "f64", C.c_double
    raise TypeError(f"unsupported dtype {dtype} (the reference instantiates float and double)")


def _p(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if hasattr(a, "data_ptr"):  # torch tensor (device pointer for the *_dev entry points)
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))


def _kernel_id(kernel):
    return _capi.KERNELS[kernel] if isinstance(kernel, str) else int(kernel)


class Context:
    """erl_gp_context: one per host thread and device."""

    def __init__(self, device: int = 0):
        self.lib = load()
        self.handle = C.c_void_p()
        check(self.lib.erl_gp_context_create(C.c_int(device), C.byref(self.handle)), "erl_gp_context_create")
        self.device = device

    def close(self):
        if getattr(self, "handle", None):
            self.lib.erl_gp_context_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int | None):
        check(self.lib.erl_gp_context_set_stream(self.handle, C.c_void_p(cuda_stream or 0)), "set_stream", self.handle)

    def synchronize(self):
        check(self.lib.erl_gp_context_synchronize(self.handle), "synchronize", self.handle)

    def set_rowgp_tc(self, on: int):
        """Fused FP32 train + predict, n <= 128: 1 = tcgen05 / TMEM kernel, 0 = mma.sync kernel, -1 = default."""
        check(self.lib.erl_gp_context_set_rowgp_tc(self.handle, C.c_int(on)), "set_rowgp_tc", self.handle)

    @property
    def kernel_launches(self) -> int:
        n = C.c_long(0)
        check(self.lib.erl_gp_context_kernel_launches(self.handle, C.byref(n)), "kernel_launches")
        return n.value

    def fn(self, name, dtype):
        return getattr(self.lib, f"{name}_{_sfx(dtype)[0]}")


_default_ctx: dict[int, Context] = {}


def default_context(device: int = 0) -> Context:
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


# ------------------------------------------------------------------------------------------
# Covariance::ComputeKtrain / ComputeKtest
# ------------------------------------------------------------------------------------------
def compute_ktrain(kernel, scale, x, var, ctx: Context | None = None):
    """K (n, n) with K[i, i] = 1 + var[i].  x: (n, x_dim)."""
    ctx = ctx or default_context()
    x = np.ascontiguousarray(x)
    n, d = x.shape
    _, ct = _sfx(x.dtype)
    var = np.ascontiguousarray(var, dtype=x.dtype)
    k = np.empty((n, n), dtype=x.dtype)
    check(ctx.fn("erl_gp_compute_ktrain", x.dtype)(ctx.handle, C.c_int(_kernel_id(kernel)), ct(scale), C.c_long(d), _p(x), C.c_long(d), _p(var), C.c_long(n), _p(k), C.c_long(n)),
          "compute_ktrain", ctx.handle)
    return k.T  # column-major n x n


def compute_ktest(kernel, scale, x1, x2, ctx: Context | None = None):
    """Ktest (n1, n2): [i, j] = k(x1_i, x2_j)."""
    ctx = ctx or default_context()
    x1 = np.ascontiguousarray(x1)
    x2 = np.ascontiguousarray(x2, dtype=x1.dtype)
    n1, d = x1.shape
    n2 = x2.shape[0]
    _, ct = _sfx(x1.dtype)
    k = np.empty((n2, n1), dtype=x1.dtype)
    check(ctx.fn("erl_gp_compute_ktest", x1.dtype)(ctx.handle, C.c_int(_kernel_id(kernel)), ct(scale), C.c_long(d), _p(x1), C.c_long(d), C.c_long(n1), _p(x2), C.c_long(d), C.c_long(n2),
                                                   _p(k), C.c_long(n1)), "compute_ktest", ctx.handle)
    return k.T


# ------------------------------------------------------------------------------------------
# VanillaGaussianProcess
# ------------------------------------------------------------------------------------------
class VanillaGaussianProcess:
    """Mirror of erl::gaussian_process::VanillaGaussianProcess<Dtype> (include/.../vanilla_gp.hpp).

    ``train(x, y, var)`` = Reset + fill TrainSet + Train (python/binding/bind_vanilla_gp.cpp:79-101);
    ``test(x_test)`` returns a TestResult with ``get_mean(y_index)`` / ``get_variance()``.
    """

    class Setting:
        def __init__(self, kernel_type="rbf", scale=1.0, max_num_samples=256):
            self.kernel_type = kernel_type
            self.scale = scale
            self.max_num_samples = max_num_samples  # vanilla_gp.hpp:28

    class TestResult:
        def __init__(self, gp, x_test):
            self._gp = gp
            self._x_test = np.ascontiguousarray(x_test, dtype=gp.dtype)
            self.num_test = self._x_test.shape[0]
            self._mean = None
            self._var = None

        def _run(self, want_mean, want_var):
            gp = self._gp
            t, d = self._x_test.shape
            mean = np.empty((gp.y_dim, t), dtype=gp.dtype) if want_mean else None
            var = np.empty(t, dtype=gp.dtype) if want_var else None
            check(gp.ctx.fn("erl_gp_vanilla_test", gp.dtype)(gp.handle, C.c_long(t), _p(self._x_test), C.c_long(d), _p(mean), _p(var)), "vanilla_test", gp.ctx.handle)
            if want_mean:
                self._mean = mean
            if want_var:
                self._var = var

        def get_mean(self, y_index=0, parallel=True):
            if self._mean is None:
                self._run(True, False)
            return self._mean[y_index].copy()

        def get_variance(self, parallel=True):
            if self._var is None:
                self._run(False, True)
            return self._var.copy()

    def __init__(self, setting: "VanillaGaussianProcess.Setting", dtype=np.float64, ctx: Context | None = None):
        self.setting = setting
        self.dtype = np.dtype(dtype)
        self.ctx = ctx or default_context()
        self.handle = C.c_void_p()
        check(self.ctx.fn("erl_gp_vanilla_create", dtype)(self.ctx.handle, C.byref(self.handle)), "vanilla_create", self.ctx.handle)
        self.is_trained = False
        self.n = 0
        self.y_dim = 1
        self.info = 0

    def __del__(self):
        try:
            if self.handle:
                self.ctx.fn("erl_gp_vanilla_destroy", self.dtype)(self.handle)
                self.handle = None
        except Exception:
            pass

    def train(self, x, y, var) -> bool:
        x = np.ascontiguousarray(x, dtype=self.dtype)
        n, d = x.shape
        s = self.setting
        if n <= 0:
            return False  # UpdateKtrain: num_samples <= 0 -> false (src/vanilla_gp.cpp:481-484)
        if not (s.max_num_samples < 0 or n <= s.max_num_samples):
            raise ValueError(f"max_num_samples should be <= {s.max_num_samples}")  # ERL_ASSERTM, src/vanilla_gp.cpp:389-392
        y = np.asarray(y, dtype=self.dtype)
        if y.ndim == 1:
            y = y[:, None]
        yf = np.asfortranarray(y)
        var = np.ascontiguousarray(var, dtype=self.dtype)
        _, ct = _sfx(self.dtype)
        info = C.c_int(0)
        check(self.ctx.fn("erl_gp_vanilla_train", self.dtype)(self.handle, C.c_int(_kernel_id(s.kernel_type)), ct(s.scale), C.c_long(d), C.c_long(y.shape[1]), C.c_long(n), _p(x),
                                                              C.c_long(d), _p(yf), C.c_long(n), _p(var), C.byref(info)), "vanilla_train", self.ctx.handle)
        self.n, self.y_dim, self.info = n, y.shape[1], info.value
        self.is_trained = True
        return True

    def replicate_to(self, other: "VanillaGaussianProcess"):
        """erl_gp_vanilla_replicate: copy the trained state (x_train, L, alpha) to `other`, a GP on another context / device."""
        check(self.ctx.fn("erl_gp_vanilla_replicate", self.dtype)(self.handle, other.handle), "vanilla_replicate", self.ctx.handle)
        other.n, other.y_dim, other.info, other.is_trained = self.n, self.y_dim, self.info, True

    @staticmethod
    def test_multi(gps, x_test):
        """erl_gp_vanilla_test_multi: one trained GP replicated on several devices, test points split into contiguous ranges, one
        host thread per replica.  Returns (mean (y_dim, T), variance (T))."""
        g0 = gps[0]
        x_test = np.ascontiguousarray(x_test, dtype=g0.dtype)
        t, d = x_test.shape
        mean = np.empty((g0.y_dim, t), dtype=g0.dtype)
        var = np.empty(t, dtype=g0.dtype)
        handles = (C.c_void_p * len(gps))(*[g.handle for g in gps])
        check(g0.ctx.fn("erl_gp_vanilla_test_multi", g0.dtype)(handles, C.c_long(len(gps)), C.c_long(t), _p(x_test), C.c_long(d), _p(mean), _p(var)), "vanilla_test_multi", g0.ctx.handle)
        return mean, var

    # ---- device-resident entry points (x, y, var, x_test, mean, var are device pointers, e.g. torch tensors) ----
    def train_dev(self, x, y, var, n, x_dim, y_dim=1):
        """erl_gp_vanilla_train_dev: asynchronous on the context's stream; `info` is read by vanilla_info()."""
        s = self.setting
        _, ct = _sfx(self.dtype)
        check(self.ctx.fn("erl_gp_vanilla_train_dev", self.dtype)(self.handle, C.c_int(_kernel_id(s.kernel_type)), ct(s.scale), C.c_long(x_dim), C.c_long(y_dim), C.c_long(n), _p(x),
                                                                  C.c_long(x_dim), _p(y), C.c_long(n), _p(var)), "vanilla_train_dev", self.ctx.handle)
        self.n, self.y_dim = n, y_dim
        self.is_trained = True

    def vanilla_info(self):
        info = C.c_int(0)
        check(self.ctx.fn("erl_gp_vanilla_info", self.dtype)(self.handle, C.byref(info)), "vanilla_info", self.ctx.handle)
        self.info = info.value
        return info.value

    def test_dev(self, x_test, num_test, x_dim, mean, var):
        check(self.ctx.fn("erl_gp_vanilla_test_dev", self.dtype)(self.handle, C.c_long(num_test), _p(x_test), C.c_long(x_dim), _p(mean), _p(var)), "vanilla_test_dev", self.ctx.handle)

    def get(self):
        """(K, L, alpha) materialised on the host, as GetKtrain / GetCholeskyDecomposition / GetAlpha."""
        n = self.n
        k = np.empty((n, n), dtype=self.dtype)
        l = np.empty((n, n), dtype=self.dtype)
        a = np.empty((self.y_dim, n), dtype=self.dtype)
        check(self.ctx.fn("erl_gp_vanilla_get", self.dtype)(self.handle, _p(k), C.c_long(n), _p(l), C.c_long(n), _p(a), C.c_long(n)), "vanilla_get", self.ctx.handle)
        return k.T, l.T, a.T

    def test(self, x_test):
        if not self.is_trained:
            return None  # src/vanilla_gp.cpp:556-558
        x_test = np.asarray(x_test)
        if x_test.shape[0] == 0:
            return None
        return VanillaGaussianProcess.TestResult(self, x_test)


# ------------------------------------------------------------------------------------------
# Batched small GPs (BatchGaussianProcessUpdateTorch replacement / BASELINE config 4)
# ------------------------------------------------------------------------------------------
class BatchGp:
    """A device-resident stream of ``num_gps`` independent small GPs with capacity ``max_n``."""

    def __init__(self, num_gps, max_n, x_dim, kernel, scale, dtype=np.float32, ctx: Context | None = None):
        self.ctx = ctx or default_context()
        self.dtype = np.dtype(dtype)
        self.num_gps, self.max_n, self.x_dim = int(num_gps), int(max_n), int(x_dim)
        _, self.ct = _sfx(dtype)
        self.handle = C.c_void_p()
        check(self.ctx.fn("erl_gp_batch_create", dtype)(self.ctx.handle, C.c_long(num_gps), C.c_long(max_n), C.c_long(x_dim), C.c_int(_kernel_id(kernel)), self.ct(scale),
                                                       C.byref(self.handle)), "batch_create", self.ctx.handle)

    def __del__(self):
        try:
            if self.handle:
                self.ctx.fn("erl_gp_batch_destroy", self.dtype)(self.handle)
                self.handle = None
        except Exception:
            pass

    # ---- host-buffer path (H2D + kernels + D2H inside; the e2e call) ----
    def train_predict(self, n_train, x, y, var, q_offsets, q_x, min_num_samples=0, want_l=True, mean=None, variance=None):
        b, mn, d = self.num_gps, self.max_n, self.x_dim
        n_train = np.ascontiguousarray(n_train, dtype=np.int32)
        x = np.ascontiguousarray(x, dtype=self.dtype).reshape(b, mn, d)
        y = np.ascontiguousarray(y, dtype=self.dtype).reshape(b, mn)
        var = np.ascontiguousarray(var, dtype=self.dtype).reshape(b, mn)
        q_offsets = np.ascontiguousarray(q_offsets, dtype=np.int64)
        q_x = np.ascontiguousarray(q_x, dtype=self.dtype).reshape(-1, d)
        t = q_x.shape[0]
        l = np.zeros((b, mn, mn), dtype=self.dtype) if want_l else None
        alpha = np.zeros((b, mn), dtype=self.dtype)
        info = np.zeros(b, dtype=np.int32)
        mean = np.full(t, np.nan, dtype=self.dtype) if mean is None else mean
        variance = np.full(t, np.nan, dtype=self.dtype) if variance is None else variance
        valid = np.zeros(t, dtype=np.uint8)
        check(self.ctx.fn("erl_gp_batch_train_predict", self.dtype)(self.handle, C.c_long(min_num_samples), _p(n_train), _p(x), _p(y), _p(var), _p(q_offsets), _p(q_x), C.c_long(t),
                                                                   _p(l), _p(alpha), _p(info), _p(mean), _p(variance), _p(valid)), "batch_train_predict", self.ctx.handle)
        return dict(L=None if l is None else l.transpose(0, 2, 1), alpha=alpha, info=info, mean=mean, var=variance, valid=valid.astype(bool))

    # ---- device-resident path ----
    def upload(self, n_train, x, y, var):
        b, mn, d = self.num_gps, self.max_n, self.x_dim
        n_train = np.ascontiguousarray(n_train, dtype=np.int32)
        x = np.ascontiguousarray(x, dtype=self.dtype).reshape(b, mn, d)
        y = np.ascontiguousarray(y, dtype=self.dtype).reshape(b, mn)
        var = np.ascontiguousarray(var, dtype=self.dtype).reshape(b, mn)
        check(self.ctx.fn("erl_gp_batch_upload", self.dtype)(self.handle, _p(n_train), _p(x), _p(y), _p(var)), "batch_upload", self.ctx.handle)
        self.ctx.synchronize()

    def device_buffers(self):
        ptrs = [C.c_void_p() for _ in range(7)]
        check(self.ctx.fn("erl_gp_batch_device_buffers", self.dtype)(self.handle, *[C.byref(p) for p in ptrs]), "batch_device_buffers", self.ctx.handle)
        names = ["n_train", "x", "y", "var", "l", "alpha", "info"]
        return {k: (p.value or 0) for k, p in zip(names, ptrs)}

    def train_dev(self, min_num_samples=0, write_l=True):
        check(self.ctx.fn("erl_gp_batch_train_dev", self.dtype)(self.handle, C.c_long(min_num_samples), C.c_int(int(write_l))), "batch_train_dev", self.ctx.handle)

    def predict_dev(self, q_offsets, q_x, num_q, mean, var, valid=None, q_out_index=None, mapping=_capi.MAPPING_NONE, mapping_scale=1.0):
        check(self.ctx.fn("erl_gp_batch_predict_dev", self.dtype)(self.handle, _p(q_offsets), _p(q_x), _p(q_out_index), C.c_long(num_q), C.c_int(mapping), self.ct(mapping_scale), _p(mean),
                                                                 _p(var), _p(valid)), "batch_predict_dev", self.ctx.handle)

    def train_predict_dev(self, q_offsets, q_x, num_q, mean, var, valid=None, min_num_samples=0, write_l=True):
        check(self.ctx.fn("erl_gp_batch_train_predict_dev", self.dtype)(self.handle, C.c_long(min_num_samples), C.c_int(int(write_l)), _p(q_offsets), _p(q_x), C.c_long(num_q), _p(mean), _p(var),
                                                                       _p(valid)), "batch_train_predict_dev", self.ctx.handle)

    def download(self, want_l=True):
        b, mn = self.num_gps, self.max_n
        l = np.zeros((b, mn, mn), dtype=self.dtype) if want_l else None
        alpha = np.zeros((b, mn), dtype=self.dtype)
        info = np.zeros(b, dtype=np.int32)
        check(self.ctx.fn("erl_gp_batch_download", self.dtype)(self.handle, _p(l), _p(alpha), _p(info)), "batch_download", self.ctx.handle)
        return dict(L=None if l is None else l.transpose(0, 2, 1), alpha=alpha, info=info)


class MultiDeviceBatchGp:
    """One GP stream over several GPUs from ONE process (erl_gp_batch_train_predict_multi_*): contiguous GP ranges per device,
    no exchange between devices, every device writes straight into the caller's arrays (src/lidar_gp_2d.cpp:366-392: the
    partitions are independent).  ``devices`` may name a device more than once (two pipelines on one GPU)."""

    def __init__(self, num_gps, max_n, x_dim, kernel, scale, dtype=np.float32, devices=None):
        devices = list(range(_capi.device_count())) if devices is None else list(devices)
        if not devices:
            raise ErlGpError(-1, "MultiDeviceBatchGp", "no CUDA device")
        self.dtype = np.dtype(dtype)
        self.num_gps, self.max_n, self.x_dim = int(num_gps), int(max_n), int(x_dim)
        nd = len(devices)
        # contiguous ranges, the first (num_gps % nd) devices take one GP more (the arithmetic of sharding.shard_range)
        self.counts = [num_gps // nd + (1 if i < num_gps % nd else 0) for i in range(nd)]
        if any(c == 0 for c in self.counts):
            raise ErlGpError(-1, "MultiDeviceBatchGp", "fewer GPs than devices")
        self.contexts = [Context(dev) for dev in devices]
        self.parts = [BatchGp(c, max_n, x_dim, kernel, scale, dtype, ctx) for c, ctx in zip(self.counts, self.contexts)]

    def train_predict(self, n_train, x, y, var, q_offsets, q_x, min_num_samples=0, want_l=False, want_alpha=True, out=None):
        """``out``: optional dict of preallocated (pinned) arrays mean / var / valid / info / alpha / L to write into."""
        b, mn, d = self.num_gps, self.max_n, self.x_dim
        n_train = np.ascontiguousarray(n_train, dtype=np.int32)
        x = np.ascontiguousarray(x, dtype=self.dtype).reshape(b, mn, d)
        y = np.ascontiguousarray(y, dtype=self.dtype).reshape(b, mn)
        var = np.ascontiguousarray(var, dtype=self.dtype).reshape(b, mn)
        q_offsets = np.ascontiguousarray(q_offsets, dtype=np.int64)
        q_x = np.ascontiguousarray(q_x, dtype=self.dtype).reshape(-1, d)
        t = q_x.shape[0]
        out = out or {}

        def buf(name, make):  # (defaults are built only when the caller did not bring a buffer)
            return out[name] if name in out else make()

        l = buf("L", lambda: np.zeros((b, mn, mn), dtype=self.dtype) if want_l else None)
        alpha = buf("alpha", lambda: np.zeros((b, mn), dtype=self.dtype) if want_alpha else None)
        info = buf("info", lambda: np.zeros(b, dtype=np.int32))
        mean = buf("mean", lambda: np.full(t, np.nan, dtype=self.dtype))
        variance = buf("var", lambda: np.full(t, np.nan, dtype=self.dtype))
        valid = buf("valid", lambda: np.zeros(t, dtype=np.uint8))
        handles = (C.c_void_p * len(self.parts))(*[part.handle for part in self.parts])
        ctx0 = self.contexts[0]
        check(ctx0.fn("erl_gp_batch_train_predict_multi", self.dtype)(handles, C.c_long(len(self.parts)), C.c_long(min_num_samples), _p(n_train), _p(x), _p(y), _p(var), _p(q_offsets),
                                                                     _p(q_x), C.c_long(t), _p(l), _p(alpha), _p(info), _p(mean), _p(variance), _p(valid)), "batch_train_predict_multi",
              ctx0.handle)
        return dict(L=None if l is None else l.transpose(0, 2, 1), alpha=alpha, info=info, mean=mean, var=variance, valid=valid if "valid" in out else valid.astype(bool))


# ------------------------------------------------------------------------------------------
# Sensor frames: minimal stand-ins for erl_geometry::LidarFrame2D / LidarFrame3D outputs.
# They only produce the arrays the hot path consumes (angles, hit / continuity masks, frame
# coordinates); erl_geometry itself is out of scope.
# ------------------------------------------------------------------------------------------
class LidarFrame2D:
    class Setting:
        def __init__(self, angle_min=-np.pi, angle_max=np.pi, num_rays=360, valid_range_min=0.0, valid_range_max=np.inf, discontinuity_detection=False,
                     discontinuity_factor=10.0, rolling_diff_discount=0.9):
            self.angle_min, self.angle_max, self.num_rays = angle_min, angle_max, num_rays
            self.valid_range_min, self.valid_range_max = valid_range_min, valid_range_max
            self.discontinuity_detection = discontinuity_detection
            self.discontinuity_factor = discontinuity_factor
            self.rolling_diff_discount = rolling_diff_discount
            self.angles = None  # explicit ray angles (e.g. a sensor log's own) instead of linspace(angle_min, angle_max, num_rays)

    def __init__(self, setting, dtype=np.float64):
        self.setting = setting
        self.dtype = np.dtype(dtype)
        if setting.angles is not None:
            self.angles = np.ascontiguousarray(setting.angles, dtype=self.dtype)
        else:
            self.angles = np.linspace(setting.angle_min, setting.angle_max, setting.num_rays).astype(self.dtype)
        self.rotation = np.eye(2, dtype=self.dtype)
        self.translation = np.zeros(2, dtype=self.dtype)
        self.ranges = None
        self.mask_hit = None
        self.mask_continuous = None

    def update_ranges(self, rotation, translation, ranges):
        s = self.setting
        self.rotation = np.asarray(rotation, dtype=self.dtype)
        self.translation = np.asarray(translation, dtype=self.dtype)
        r = np.asarray(ranges, dtype=self.dtype)
        self.ranges = r
        self.mask_hit = np.isfinite(r) & (r >= s.valid_range_min) & (r <= s.valid_range_max)
        con = np.ones(len(r), dtype=bool)
        if s.discontinuity_detection:
            # rolling mean of |range difference| between consecutive rays; a jump larger than
            # discontinuity_factor x the rolling mean marks both rays as discontinuous
            rolling = 0.0
            for i in range(1, len(r)):
                diff = abs(float(r[i]) - float(r[i - 1]))
                if i > 1 and diff > s.discontinuity_factor * rolling and rolling > 0:
                    con[i - 1] = con[i] = False
                rolling = s.rolling_diff_discount * rolling + (1 - s.rolling_diff_discount) * diff if i > 1 else diff
        self.mask_continuous = con

    @property
    def is_valid(self):
        return self.mask_hit is not None and bool(self.mask_hit.any())


class LidarGaussianProcess2D:
    """Mirror of erl::gaussian_process::LidarGaussianProcess2D<Dtype> (include/.../lidar_gp_2d.hpp)."""

    class Setting:
        def __init__(self):
            # defaults: include/erl_gaussian_process/lidar_gp_2d.hpp:28-62
            self.partition_on_hit_rays = False
            self.symmetric_partitions = True
            self.group_size = 26
            self.overlap_size = 6
            self.margin = 1
            self.init_variance = 1e6
            self.sensor_range_var = 0.01
            self.discontinuity_var = 10.0
            self.max_valid_range_var = 0.1
            self.occ_test_temperature = 30.0
            self.sensor_frame = LidarFrame2D.Setting()
            self.gp = VanillaGaussianProcess.Setting(kernel_type="ou", scale=1.0)
            self.mapping_type = _capi.MAPPING_INVERSE_SQRT
            self.mapping_scale = 1.0

    class _CSetting(C.Structure):
        _fields_ = [("symmetric_partitions", C.c_int), ("group_size", C.c_long), ("overlap_size", C.c_long), ("margin", C.c_long), ("sensor_range_var", C.c_double),
                    ("discontinuity_var", C.c_double), ("discontinuity_detection", C.c_int), ("kernel", C.c_int), ("kernel_scale", C.c_double), ("mapping", C.c_int),
                    ("mapping_scale", C.c_double), ("partition_on_hit_rays", C.c_int)]

    class TestResult:
        def __init__(self, gp, angles, angles_are_local, un_map):
            angles = np.ascontiguousarray(angles, dtype=gp.dtype)
            t = len(angles)
            self.num_test = t
            # invalid rays are left unwritten by the reference (src/lidar_gp_2d.cpp:112,120): NaN marks them here
            self._mean = np.full(t, np.nan, dtype=gp.dtype)
            self._var = np.full(t, np.nan, dtype=gp.dtype)
            valid = np.zeros(t, dtype=np.uint8)
            check(gp.ctx.fn("erl_gp_lidar2d_test", gp.dtype)(gp.handle, _p(angles), C.c_long(t), C.c_int(int(angles_are_local)), C.c_int(int(un_map)), _p(self._mean), _p(self._var),
                                                            _p(valid)), "lidar2d_test", gp.ctx.handle)
            self._valid = valid.astype(bool)

        def get_mean(self, parallel=True):
            return self._mean.copy(), self._valid.copy()

        def get_variance(self, parallel=True):
            return self._var.copy(), self._valid.copy()

    def __init__(self, setting: "LidarGaussianProcess2D.Setting", dtype=np.float64, ctx: Context | None = None):
        self.setting = setting
        self.dtype = np.dtype(dtype)
        self.ctx = ctx or default_context()
        self.sensor_frame = LidarFrame2D(setting.sensor_frame, dtype)
        cs = self._CSetting(int(setting.symmetric_partitions), setting.group_size, setting.overlap_size, setting.margin, setting.sensor_range_var, setting.discontinuity_var,
                            int(setting.sensor_frame.discontinuity_detection), _kernel_id(setting.gp.kernel_type), setting.gp.scale, setting.mapping_type, setting.mapping_scale,
                            int(setting.partition_on_hit_rays))  # src/lidar_gp_2d.cpp:302-348 with the out-of-range indices clamped
        self.handle = C.c_void_p()
        angles = self.sensor_frame.angles
        check(self.ctx.fn("erl_gp_lidar2d_create", dtype)(self.ctx.handle, C.byref(cs), _p(angles), C.c_long(len(angles)), C.byref(self.handle)), "lidar2d_create", self.ctx.handle)
        self.is_trained = False

    def __del__(self):
        try:
            if self.handle:
                self.ctx.fn("erl_gp_lidar2d_destroy", self.dtype)(self.handle)
                self.handle = None
        except Exception:
            pass

    @property
    def num_partitions(self):
        n = C.c_long(0)
        check(self.ctx.fn("erl_gp_lidar2d_num_partitions", self.dtype)(self.handle, C.byref(n)), "lidar2d_num_partitions")
        return n.value

    @property
    def angle_partitions(self):
        n = self.num_partitions
        il, ir = np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.int64)
        cl, cr = np.zeros(n, dtype=self.dtype), np.zeros(n, dtype=self.dtype)
        check(self.ctx.fn("erl_gp_lidar2d_partitions", self.dtype)(self.handle, _p(il), _p(ir), _p(cl), _p(cr)), "lidar2d_partitions")
        return [(int(il[i]), int(ir[i]), cl[i], cr[i]) for i in range(n)]

    def train(self, rotation, translation, ranges) -> bool:
        self.is_trained = False
        frame = self.sensor_frame
        frame.update_ranges(rotation, translation, ranges)
        if not frame.is_valid:
            return False
        rot = np.asfortranarray(frame.rotation)
        r = np.ascontiguousarray(frame.ranges)
        hit = np.ascontiguousarray(frame.mask_hit, dtype=np.uint8)
        con = np.ascontiguousarray(frame.mask_continuous, dtype=np.uint8)
        check(self.ctx.fn("erl_gp_lidar2d_train", self.dtype)(self.handle, _p(rot), _p(r), _p(hit), _p(con)), "lidar2d_train", self.ctx.handle)
        self.is_trained = True
        return True

    def test(self, angles, angles_are_local, un_map=True):
        if not self.is_trained:
            return None
        return LidarGaussianProcess2D.TestResult(self, angles, angles_are_local, un_map)

    def get_gp(self, p):
        gs = self.setting.group_size
        info, n = C.c_int(0), C.c_long(0)
        l = np.zeros((gs, gs), dtype=self.dtype)
        a = np.zeros(gs, dtype=self.dtype)
        check(self.ctx.fn("erl_gp_lidar2d_get_gp", self.dtype)(self.handle, C.c_long(p), C.byref(info), C.byref(n), _p(l), C.c_long(gs), _p(a)), "lidar2d_get_gp", self.ctx.handle)
        nn = n.value
        return info.value, nn, l.T[:nn, :nn].copy(), a[:nn].copy()

    def compute_occ(self, pos_local):
        """Batched ComputeOcc (src/lidar_gp_2d.cpp:428-459). pos_local: (T, 2). Returns ok, dist, range_pred, occ."""
        pos = np.ascontiguousarray(pos_local, dtype=self.dtype)
        t = pos.shape[0]
        _, ct = _sfx(self.dtype)
        dist = np.full(t, np.nan, dtype=self.dtype)
        rp = np.full(t, np.nan, dtype=self.dtype)
        occ = np.full(t, np.nan, dtype=self.dtype)
        ok = np.zeros(t, dtype=np.uint8)
        check(self.ctx.fn("erl_gp_lidar2d_compute_occ", self.dtype)(self.handle, _p(pos), C.c_long(t), ct(self.setting.max_valid_range_var), ct(self.setting.occ_test_temperature), _p(dist),
                                                                   _p(rp), _p(occ), _p(ok)), "lidar2d_compute_occ", self.ctx.handle)
        return ok.astype(bool), dist, rp, occ


class LidarFrame3D:
    """Azimuth x elevation ray grid: frame_coords[r, c] = (azimuth_r, elevation_c)
    (test/gtest/test_range_sensor_gp_3d.cpp:39-44; rows follow azimuth, cols elevation)."""

    class Setting:
        def __init__(self, azimuth_min=-np.pi, azimuth_max=np.pi, num_azimuth_lines=360, elevation_min=-np.pi / 2, elevation_max=np.pi / 2, num_elevation_lines=181,
                     valid_range_min=0.0, valid_range_max=np.inf):
            self.azimuth_min, self.azimuth_max, self.num_azimuth_lines = azimuth_min, azimuth_max, num_azimuth_lines
            self.elevation_min, self.elevation_max, self.num_elevation_lines = elevation_min, elevation_max, num_elevation_lines
            self.valid_range_min, self.valid_range_max = valid_range_min, valid_range_max

    def __init__(self, setting, dtype=np.float32):
        self.setting = setting
        self.dtype = np.dtype(dtype)
        az = np.linspace(setting.azimuth_min, setting.azimuth_max, setting.num_azimuth_lines)
        el = np.linspace(setting.elevation_min, setting.elevation_max, setting.num_elevation_lines)
        self.frame_coords = np.stack(np.meshgrid(az, el, indexing="ij"), axis=-1).astype(self.dtype)  # (rows, cols, 2)
        self.rotation = np.eye(3, dtype=self.dtype)
        self.ranges = None
        self.mask_hit = None

    def update_ranges(self, rotation, translation, ranges):
        s = self.setting
        self.rotation = np.asarray(rotation, dtype=self.dtype)
        r = np.asarray(ranges, dtype=self.dtype)
        self.ranges = r
        self.mask_hit = np.isfinite(r) & (r >= s.valid_range_min) & (r <= s.valid_range_max)

    @property
    def is_valid(self):
        return self.mask_hit is not None and bool(self.mask_hit.any())

    def dir_world_to_frame(self, dirs):
        return np.asarray(dirs, dtype=self.dtype) @ self.rotation  # (R^T d) for row vectors

    def compute_frame_coords(self, dirs_local):
        d = np.asarray(dirs_local, dtype=self.dtype)
        dist = np.linalg.norm(d, axis=1)
        ok = dist > 0
        safe = np.where(ok, dist, 1).astype(self.dtype)
        az = np.arctan2(d[:, 1], d[:, 0])
        el = np.arcsin(np.clip(d[:, 2] / safe, -1, 1))
        return ok, dist.astype(self.dtype), np.stack([az, el], axis=1).astype(self.dtype)


class RangeSensorGaussianProcess3D:
    """Mirror of erl::gaussian_process::RangeSensorGaussianProcess3D<Dtype> (include/.../range_sensor_gp_3d.hpp)."""

    class Setting:
        def __init__(self):
            # defaults: include/erl_gaussian_process/range_sensor_gp_3d.hpp:31-74
            self.row_group_size, self.row_overlap_size, self.row_margin = 24, 6, 0
            self.col_group_size, self.col_overlap_size, self.col_margin = 8, 2, 0
            self.min_num_samples_per_group = 32
            self.init_variance = 1e6
            self.sensor_range_var = 0.01
            self.max_valid_range_var = 0.1
            self.occ_test_temperature = 30.0
            self.sensor_frame = LidarFrame3D.Setting()
            self.gp = VanillaGaussianProcess.Setting(kernel_type="ou", scale=1.0)
            self.mapping_type = _capi.MAPPING_INVERSE_SQRT
            self.mapping_scale = 1.0

    class _CSetting(C.Structure):
        _fields_ = [("row_group_size", C.c_long), ("row_overlap_size", C.c_long), ("row_margin", C.c_long), ("col_group_size", C.c_long), ("col_overlap_size", C.c_long),
                    ("col_margin", C.c_long), ("min_num_samples_per_group", C.c_long), ("sensor_range_var", C.c_double), ("kernel", C.c_int), ("kernel_scale", C.c_double),
                    ("mapping", C.c_int), ("mapping_scale", C.c_double)]

    class TestResult:
        def __init__(self, gp, directions, directions_are_local, un_map):
            frame = gp.sensor_frame
            d = np.asarray(directions, dtype=gp.dtype)
            if not directions_are_local:
                d = frame.dir_world_to_frame(d)
            ok, _, coords = frame.compute_frame_coords(d)
            self._init_from_coords(gp, coords, ok, un_map)

        def _init_from_coords(self, gp, coords, ok, un_map):
            coords = np.ascontiguousarray(coords, dtype=gp.dtype)
            t = coords.shape[0]
            self.num_test = t
            self._mean = np.full(t, np.nan, dtype=gp.dtype)
            self._var = np.full(t, np.nan, dtype=gp.dtype)
            valid = np.zeros(t, dtype=np.uint8)
            okb = None if ok is None else np.ascontiguousarray(ok, dtype=np.uint8)
            check(gp.ctx.fn("erl_gp_range3d_test", gp.dtype)(gp.handle, _p(coords), _p(okb), C.c_long(t), C.c_int(int(un_map)), _p(self._mean), _p(self._var), _p(valid)), "range3d_test",
                  gp.ctx.handle)
            self._valid = valid.astype(bool)

        def get_mean(self, parallel=True):
            return self._mean.copy(), self._valid.copy()

        def get_variance(self, parallel=True):
            return self._var.copy(), self._valid.copy()

    def __init__(self, setting: "RangeSensorGaussianProcess3D.Setting", dtype=np.float32, ctx: Context | None = None, sensor_frame=None):
        if setting.row_overlap_size % 2 or setting.col_overlap_size % 2:
            raise ValueError("row_overlap_size / col_overlap_size must be even")  # ERL_ASSERTM src/range_sensor_gp_3d.cpp:190-197
        self.setting = setting
        self.dtype = np.dtype(dtype)
        self.ctx = ctx or default_context()
        self.sensor_frame = sensor_frame or LidarFrame3D(setting.sensor_frame, dtype)
        fc = self.sensor_frame.frame_coords
        self.rows, self.cols = fc.shape[:2]
        fcc = np.ascontiguousarray(fc.transpose(1, 0, 2))  # Eigen col-major matrix of Vector2
        cs = self._CSetting(setting.row_group_size, setting.row_overlap_size, setting.row_margin, setting.col_group_size, setting.col_overlap_size, setting.col_margin,
                            setting.min_num_samples_per_group, setting.sensor_range_var, _kernel_id(setting.gp.kernel_type), setting.gp.scale, setting.mapping_type, setting.mapping_scale)
        self.handle = C.c_void_p()
        check(self.ctx.fn("erl_gp_range3d_create", dtype)(self.ctx.handle, C.byref(cs), _p(fcc), C.c_long(self.rows), C.c_long(self.cols), C.byref(self.handle)), "range3d_create",
              self.ctx.handle)
        self.is_trained = False

    def __del__(self):
        try:
            if self.handle:
                self.ctx.fn("erl_gp_range3d_destroy", self.dtype)(self.handle)
                self.handle = None
        except Exception:
            pass

    @property
    def grid(self):
        a, b = C.c_long(0), C.c_long(0)
        check(self.ctx.fn("erl_gp_range3d_grid", self.dtype)(self.handle, C.byref(a), C.byref(b)), "range3d_grid")
        return a.value, b.value

    def partitions(self, axis):
        n = self.grid[axis]
        il, ir = np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.int64)
        cl, cr = np.zeros(n, dtype=self.dtype), np.zeros(n, dtype=self.dtype)
        check(self.ctx.fn("erl_gp_range3d_partitions", self.dtype)(self.handle, C.c_int(axis), _p(il), _p(ir), _p(cl), _p(cr)), "range3d_partitions")
        return [(int(il[i]), int(ir[i]), cl[i], cr[i]) for i in range(n)]

    def train(self, rotation, translation, ranges) -> bool:
        self.is_trained = False
        frame = self.sensor_frame
        frame.update_ranges(rotation, translation, ranges)
        if not frame.is_valid:
            return False
        r = np.asfortranarray(frame.ranges)
        m = np.asfortranarray(frame.mask_hit.astype(np.uint8))
        check(self.ctx.fn("erl_gp_range3d_train", self.dtype)(self.handle, _p(r), _p(m)), "range3d_train", self.ctx.handle)
        self.is_trained = True
        return True

    def test(self, directions, directions_are_local, un_map=True):
        if not self.is_trained:
            return None
        return RangeSensorGaussianProcess3D.TestResult(self, directions, directions_are_local, un_map)

    def test_frame_coords(self, coords, coords_ok=None, un_map=True):
        if not self.is_trained:
            return None
        res = RangeSensorGaussianProcess3D.TestResult.__new__(RangeSensorGaussianProcess3D.TestResult)
        res._init_from_coords(self, coords, coords_ok, un_map)
        return res

    def compute_occ(self, pos_local):
        """Batched ComputeOcc (src/range_sensor_gp_3d.cpp:409-439). pos_local: (T, 3) in the sensor frame.
        Returns ok, dist, range_pred, occ (range_pred / occ of rejected positions stay NaN)."""
        if not self.is_trained:
            return None
        pos = np.asarray(pos_local, dtype=self.dtype)
        ok_c, dist, coords = self.sensor_frame.compute_frame_coords(pos)  # ComputeFrameCoords + CoordsIsInFrame stay on the host side
        t = pos.shape[0]
        _, ct = _sfx(self.dtype)
        coords = np.ascontiguousarray(coords, dtype=self.dtype)
        dist = np.ascontiguousarray(dist, dtype=self.dtype)
        okb = np.ascontiguousarray(ok_c, dtype=np.uint8)
        rp = np.full(t, np.nan, dtype=self.dtype)
        occ = np.full(t, np.nan, dtype=self.dtype)
        ok = np.zeros(t, dtype=np.uint8)
        check(self.ctx.fn("erl_gp_range3d_compute_occ", self.dtype)(self.handle, _p(coords), _p(okb), _p(dist), C.c_long(t), ct(self.setting.max_valid_range_var),
                                                                   ct(self.setting.occ_test_temperature), _p(rp), _p(occ), _p(ok)), "range3d_compute_occ", self.ctx.handle)
        return ok.astype(bool), dist, rp, occ

    def get_gp(self, row_part, col_part):
        mn = self.setting.row_group_size * self.setting.col_group_size
        info, n = C.c_int(0), C.c_long(0)
        l = np.zeros((mn, mn), dtype=self.dtype)
        a = np.zeros(mn, dtype=self.dtype)
        check(self.ctx.fn("erl_gp_range3d_get_gp", self.dtype)(self.handle, C.c_long(row_part), C.c_long(col_part), C.byref(info), C.byref(n), _p(l), C.c_long(mn), _p(a)), "range3d_get_gp",
              self.ctx.handle)
        nn = n.value
        return info.value, nn, l.T[:nn, :nn].copy(), a[:nn].copy()


class SparsePseudoInputGaussianProcess:
    """Mirror of erl::gaussian_process::SparsePseudoInputGaussianProcess<Dtype>, dense mode."""

    def __init__(self, kernel, scale, pseudo_points, dtype=np.float64, ctx: Context | None = None):
        self.ctx = ctx or default_context()
        self.dtype = np.dtype(dtype)
        _, ct = _sfx(dtype)
        z = np.ascontiguousarray(pseudo_points, dtype=self.dtype)
        self.m, self.x_dim = z.shape
        self.handle = C.c_void_p()
        check(self.ctx.fn("erl_gp_spgp_create", dtype)(self.ctx.handle, C.c_int(_kernel_id(kernel)), ct(scale), C.c_long(self.x_dim), C.c_long(self.m), _p(z), C.byref(self.handle)),
              "spgp_create", self.ctx.handle)

    def __del__(self):
        try:
            if self.handle:
                self.ctx.fn("erl_gp_spgp_destroy", self.dtype)(self.handle)
                self.handle = None
        except Exception:
            pass

    def update(self, x, y, var) -> bool:
        x = np.ascontiguousarray(x, dtype=self.dtype)
        n = x.shape[0]
        if n == 0:
            return False  # src/sparse_pseudo_input_gp.cpp:755
        y = np.ascontiguousarray(y, dtype=self.dtype)
        var = np.ascontiguousarray(var, dtype=self.dtype)
        check(self.ctx.fn("erl_gp_spgp_update", self.dtype)(self.handle, C.c_long(n), _p(x), C.c_long(self.x_dim), _p(y), _p(var)), "spgp_update", self.ctx.handle)
        return True

    def test(self, x_test):
        xt = np.ascontiguousarray(x_test, dtype=self.dtype)
        t = xt.shape[0]
        mean = np.empty(t, dtype=self.dtype)
        var = np.empty(t, dtype=self.dtype)
        check(self.ctx.fn("erl_gp_spgp_test", self.dtype)(self.handle, C.c_long(t), _p(xt), C.c_long(self.x_dim), _p(mean), _p(var)), "spgp_test", self.ctx.handle)
        return mean, var

    def set_diagonal_qm(self, on=True):
        """Setting::diagonal_qm: Q_M kept as its diagonal (call before the first update); mean and gradient only."""
        check(self.ctx.fn("erl_gp_spgp_set_diagonal_qm", self.dtype)(self.handle, C.c_int(int(on))), "spgp_set_diagonal_qm", self.ctx.handle)

    def get_qm_diagonal(self):
        q = np.empty(self.m, dtype=self.dtype)
        check(self.ctx.fn("erl_gp_spgp_get_qm_diagonal", self.dtype)(self.handle, _p(q)), "spgp_get_qm_diagonal", self.ctx.handle)
        return q

    def test_mean(self, x_test):
        xt = np.ascontiguousarray(x_test, dtype=self.dtype)
        t = xt.shape[0]
        mean = np.empty(t, dtype=self.dtype)
        check(self.ctx.fn("erl_gp_spgp_test", self.dtype)(self.handle, C.c_long(t), _p(xt), C.c_long(self.x_dim), _p(mean), None), "spgp_test", self.ctx.handle)
        return mean

    def test_gradient(self, x_test, raw_alpha=False):
        """TestResult::GetGradient (src/sparse_pseudo_input_gp.cpp:187-278): gradient of the predictive mean, (T, x_dim).
        raw_alpha=True dots with the unsolved alpha, as the reference's batched accessor does (:212)."""
        xt = np.ascontiguousarray(x_test, dtype=self.dtype)
        t = xt.shape[0]
        grad = np.empty((t, self.x_dim), dtype=self.dtype)
        check(self.ctx.fn("erl_gp_spgp_test_gradient", self.dtype)(self.handle, C.c_long(t), _p(xt), C.c_long(self.x_dim), _p(grad), C.c_int(int(raw_alpha))), "spgp_test_gradient",
              self.ctx.handle)
        return grad

    def set_state(self, q_m, alpha):
        """Read() of the reference (src/sparse_pseudo_input_gp.cpp:721-740): restores the accumulated Q_M ((M, M); (M,) in diagonal_qm
        mode) and alpha (M,) into an instance created with the same pseudo-points; L_QM is refactored at the next test."""
        q = np.asarray(q_m, dtype=self.dtype)
        q = np.ascontiguousarray(q.T if q.ndim == 2 else q)  # the C ABI is column-major (Q_M is symmetric up to rounding: keep the caller's entries)
        a = np.ascontiguousarray(alpha, dtype=self.dtype)
        if a.shape != (self.m,) or q.shape not in ((self.m, self.m), (self.m,)):
            raise ValueError("set_state: q_m must be (M, M) or (M,), alpha (M,)")
        check(self.ctx.fn("erl_gp_spgp_set_state", self.dtype)(self.handle, _p(q), _p(a)), "spgp_set_state", self.ctx.handle)

    def get(self):
        m = self.m
        q = np.empty((m, m), dtype=self.dtype)
        a = np.empty(m, dtype=self.dtype)
        lk = np.empty((m, m), dtype=self.dtype)
        lq = np.empty((m, m), dtype=self.dtype)
        check(self.ctx.fn("erl_gp_spgp_get", self.dtype)(self.handle, _p(q), _p(a), _p(lk), _p(lq)), "spgp_get", self.ctx.handle)
        return q.T, a, lk.T, lq.T


class SpGpOccupancyMap:
    """Mirror of erl::gaussian_process::SpGpOccupancyMap<Dtype, Dim> (include/.../spgp_occupancy_map.hpp, src/spgp_occupancy_map.cpp):
    a log-odds occupancy field regressed by an SPGP.  ``update_with_dataset`` is Update() after the dataset exists (:106-123: labels ->
    logodd_occupied / logodd_free, constant logodd_variance, SPGP Reset + Update); ``predict`` is Predict (:126-140).  The dataset
    generator of the reference lives in erl_geometry (absent dependency): ``generate_dataset`` is a stand-in with the documented
    meaning of the Setting fields, its random stream is its own."""

    def __init__(self, kernel, scale, pseudo_points, boundary_center, boundary_half_sizes, seed=0, dtype=np.float64, ctx: Context | None = None,
                 min_distance=0.5, max_distance=30.0, free_points_per_meter=2.0, free_sampling_margin=0.05, logodd_free=-5.0, logodd_occupied=5.0,
                 logodd_variance=1e-4, max_num_samples=256):
        self.sp_gp = SparsePseudoInputGaussianProcess(kernel, scale, pseudo_points, dtype, ctx)
        self.dtype = np.dtype(dtype)
        self.dim = self.sp_gp.x_dim
        self.center = np.asarray(boundary_center, dtype=np.float64)
        self.half_sizes = np.asarray(boundary_half_sizes, dtype=np.float64)
        self.rng = np.random.default_rng(seed)
        self.min_distance, self.max_distance = float(min_distance), float(max_distance)
        self.free_points_per_meter, self.free_sampling_margin = float(free_points_per_meter), float(free_sampling_margin)
        self.logodd_free, self.logodd_occupied, self.logodd_variance = float(logodd_free), float(logodd_occupied), float(logodd_variance)
        self.max_num_samples = int(max_num_samples)
        self.trained = False

    def _inside(self, p):
        return np.all(np.abs(p - self.center) <= self.half_sizes, axis=-1)

    def generate_dataset(self, sensor_position, points, point_indices=None, max_dataset_size=-1):
        """-> (dataset_points (n, dim), dataset_labels (n,), hit_indices): hit points (label 1) + free points along the rays (label 0)."""
        sensor = np.asarray(sensor_position, dtype=np.float64)
        pts = np.asarray(points, dtype=np.float64)
        idx = np.arange(len(pts)) if point_indices is None or len(point_indices) == 0 else np.asarray(point_indices, dtype=np.int64)
        out_p, out_l, hits = [], [], []
        count = 0
        for i in idx:
            if 0 < max_dataset_size <= count:
                break
            v = pts[i] - sensor
            r = float(np.linalg.norm(v))
            if not np.isfinite(r) or r < self.min_distance or r > self.max_distance:
                continue
            if self._inside(pts[i]):
                out_p.append(pts[i]), out_l.append(1.0), hits.append(int(i))
                count += 1
            for _ in range(int(np.floor(r * self.free_points_per_meter))):
                if 0 < max_dataset_size <= count:
                    break
                q = sensor + self.rng.uniform(self.free_sampling_margin, 1.0 - self.free_sampling_margin) * v
                if self._inside(q):
                    out_p.append(q), out_l.append(0.0)
                    count += 1
        return np.asarray(out_p, dtype=self.dtype).reshape(-1, self.dim), np.asarray(out_l, dtype=self.dtype), hits

    def update_with_dataset(self, dataset_points, dataset_labels) -> bool:
        x = np.ascontiguousarray(dataset_points, dtype=self.dtype)
        if len(x) == 0:
            return False  # "No valid points generated for update. Skipping update."
        if len(x) > self.max_num_samples:
            raise ValueError(f"max_num_samples should be <= {self.max_num_samples}")  # SPGP Reset, src/sparse_pseudo_input_gp.cpp:405-408
        y = np.where(np.asarray(dataset_labels) > 0, self.logodd_occupied, self.logodd_free).astype(self.dtype)
        ok = self.sp_gp.update(x, y, np.full(len(x), self.logodd_variance, dtype=self.dtype))
        self.trained = self.trained or ok
        return ok

    def update(self, sensor_position, points, point_indices=None):
        """Update (:82-124) -> (ok, dataset_points, dataset_labels, hit_indices)."""
        p, l, hits = self.generate_dataset(sensor_position, points, point_indices, self.max_num_samples)
        return self.update_with_dataset(p, l), p, l, hits

    def predict(self, points, compute_gradient=False):
        """-> logodd (T,) [, gradient (T, dim)]"""
        if not self.trained:
            raise RuntimeError("predict() before the first successful update (Test() returns nullptr until trained)")
        logodd = self.sp_gp.test_mean(points)
        if not compute_gradient:
            return logodd
        return logodd, self.sp_gp.test_gradient(points)

    def predict_gradient(self, points):
        return self.predict(points, True)[1]


class NoisyInputGaussianProcess:
    """Mirror of erl::gaussian_process::NoisyInputGaussianProcess<Dtype> (include/.../noisy_input_gp.hpp): GP with noisy
    inputs and gradient observations.  ``train(x, y, grad, var_x, var_y, var_grad, grad_flag)`` = Reset + TrainSet fill +
    Train (src/noisy_input_gp.cpp:700-724, 807-899); ``test(x_test, predict_gradient)`` returns a TestResult with
    ``get_mean`` / ``get_gradient`` / ``get_mean_variance`` / ``get_gradient_variance`` / ``get_covariance`` (:125-333)."""

    class Setting:
        def __init__(self, kernel_type="rbf", scale=1.0, max_num_samples=-1, no_gradient_observation=False):
            self.kernel_type = kernel_type
            self.scale = scale
            self.max_num_samples = max_num_samples  # noisy_input_gp.hpp:24: -1 = no limit
            self.no_gradient_observation = no_gradient_observation

    class TestResult:
        def __init__(self, gp, x_test, predict_gradient):
            self._gp = gp
            self._x_test = np.ascontiguousarray(x_test, dtype=gp.dtype)
            self.num_test, self.x_dim = self._x_test.shape
            self.support_gradient = bool(predict_gradient)
            self._mean = self._grad = self._var = self._gvar = self._cov = None

        def _run(self, mean=False, var=False):
            gp, t, d = self._gp, self.num_test, self.x_dim
            sg = self.support_gradient
            m = np.empty((gp.y_dim, t), dtype=gp.dtype) if mean else None
            g = np.empty((gp.y_dim, t, d), dtype=gp.dtype) if mean and sg else None
            v = np.empty(t, dtype=gp.dtype) if var else None
            gv = np.empty((t, d), dtype=gp.dtype) if var and sg else None
            cv = np.empty((t, d * (d + 1) // 2), dtype=gp.dtype) if var and sg else None
            check(gp.ctx.fn("erl_gp_noisy_test", gp.dtype)(gp.handle, C.c_long(t), _p(self._x_test), C.c_long(d), C.c_int(int(sg)), _p(m), _p(g), _p(v), _p(gv), _p(cv)), "noisy_test",
                  gp.ctx.handle)
            if mean:
                self._mean, self._grad = m, g
            if var:
                self._var, self._gvar, self._cov = v, gv, cv

        def get_mean(self, y_index=0, parallel=True):
            if self._mean is None:
                self._run(mean=True)
            return self._mean[y_index].copy()

        def get_gradient(self, y_index=0, parallel=True):
            """-> (gradient (T, x_dim), valid (T,)): a gradient with a non-finite component is invalid (:193-197)"""
            assert self.support_gradient, "m_support_gradient_ = false"
            if self._mean is None:
                self._run(mean=True)
            g = self._grad[y_index].copy()
            return g, np.isfinite(g).all(axis=1)

        def get_mean_variance(self, parallel=True):
            if self._var is None:
                self._run(var=True)
            return self._var.copy()

        def get_gradient_variance(self, parallel=True):
            assert self.support_gradient, "m_support_gradient_ = false"
            if self._var is None:
                self._run(var=True)
            return self._gvar.copy()

        def get_covariance(self, parallel=True):
            assert self.support_gradient, "m_support_gradient_ = false"
            if self._var is None:
                self._run(var=True)
            return self._cov.copy()

    def __init__(self, setting: "NoisyInputGaussianProcess.Setting", dtype=np.float64, ctx: Context | None = None):
        self.setting = setting
        self.dtype = np.dtype(dtype)
        self.ctx = ctx or default_context()
        self.handle = C.c_void_p()
        check(self.ctx.fn("erl_gp_noisy_create", dtype)(self.ctx.handle, C.byref(self.handle)), "noisy_create", self.ctx.handle)
        self.is_trained = False
        self.n = self.m = 0
        self.y_dim = 1

    def __del__(self):
        try:
            if self.handle:
                self.ctx.fn("erl_gp_noisy_destroy", self.dtype)(self.handle)
                self.handle = None
        except Exception:
            pass

    def train(self, x, y, grad, var_x, var_y, var_grad, grad_flag) -> bool:
        """x (n, x_dim); y (n,) or (n, y_dim); grad (n, y_dim, x_dim) or None; var_*: scalars or (n,); grad_flag: scalar or (n,)"""
        x = np.ascontiguousarray(x, dtype=self.dtype)
        n, d = x.shape
        s = self.setting
        if n <= 0:
            return False  # :811-814
        if not (s.max_num_samples < 0 or n <= s.max_num_samples):
            raise ValueError(f"max_num_samples should be <= {s.max_num_samples}")  # :711-714
        y = np.asarray(y, dtype=self.dtype).reshape(n, -1)
        y_dim = y.shape[1]
        yf = np.asfortranarray(y)
        g = None if grad is None else np.ascontiguousarray(np.asarray(grad, dtype=self.dtype).reshape(n, y_dim * d))
        vx, vy = (np.ascontiguousarray(np.broadcast_to(v, (n,)), dtype=self.dtype) for v in (var_x, var_y))
        vg = None if var_grad is None else np.ascontiguousarray(np.broadcast_to(var_grad, (n,)), dtype=self.dtype)
        flag = np.ascontiguousarray(np.broadcast_to(grad_flag, (n,)), dtype=np.int64)
        _, ct = _sfx(self.dtype)
        check(self.ctx.fn("erl_gp_noisy_train", self.dtype)(self.handle, C.c_int(_kernel_id(s.kernel_type)), ct(s.scale), C.c_long(d), C.c_long(y_dim), C.c_long(n), _p(x), C.c_long(d), _p(yf),
                                                            C.c_long(n), _p(g), C.c_long(y_dim * d), _p(vx), _p(vy), _p(vg), _p(flag), C.c_int(int(s.no_gradient_observation))), "noisy_train",
              self.ctx.handle)
        self.n, self.y_dim = n, y_dim
        self.is_trained = True
        return True

    def test(self, x_test, predict_gradient=True):
        if not self.is_trained:
            return None  # :905
        return NoisyInputGaussianProcess.TestResult(self, x_test, predict_gradient)

    def get(self):
        """-> info, K (m, m), L (m, m), alpha (m, y_dim): GetKtrainSized / GetCholeskyDecomposition / GetAlphaSized"""
        rows, info = C.c_long(0), C.c_int(0)
        check(self.ctx.fn("erl_gp_noisy_get", self.dtype)(self.handle, C.byref(rows), C.byref(info), None, C.c_long(0), None, C.c_long(0), None, C.c_long(0)), "noisy_get", self.ctx.handle)
        m = self.m = rows.value
        k, l = np.zeros((m, m), dtype=self.dtype), np.zeros((m, m), dtype=self.dtype)
        a = np.zeros((self.y_dim, m), dtype=self.dtype)
        check(self.ctx.fn("erl_gp_noisy_get", self.dtype)(self.handle, C.byref(rows), C.byref(info), _p(k), C.c_long(m), _p(l), C.c_long(m), _p(a), C.c_long(m)), "noisy_get", self.ctx.handle)
        return info.value, k.T.copy(), l.T.copy(), a.T.copy()
