"""The pybind11 module (erl_gaussian_process_b200/python/binding): import and surface on CPU, numerics on the GPU.
Mirrors the reference's pyerl_gaussian_process (python/binding/*.cpp): class names, nested Setting / TestResult, method and
argument names."""
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(ROOT, "erl_gaussian_process_b200", "lib")


@pytest.fixture(scope="module")
def pygp():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "erl_gaussian_process_b200", "python", "binding"), f"PY={sys.executable}"], stdout=subprocess.DEVNULL)
    if LIB not in sys.path:
        sys.path.insert(0, LIB)
    import pyerl_gaussian_process_b200 as m

    return m


def test_module_surface(pygp):
    for name in ("VanillaGaussianProcessD", "VanillaGaussianProcessF", "NoisyInputGaussianProcessD", "NoisyInputGaussianProcessF", "LidarGaussianProcess2Dd", "LidarGaussianProcess2Df",
                 "RangeSensorGaussianProcess3Dd", "RangeSensorGaussianProcess3Df", "MappingD", "MappingType"):
        assert hasattr(pygp, name), name
    s = pygp.VanillaGaussianProcessD.Setting()
    assert s.max_num_samples == 256 and s.kernel.x_dim == -1                     # vanilla_gp.hpp:28
    ls = pygp.LidarGaussianProcess2Df.Setting()
    assert (ls.group_size, ls.overlap_size, ls.margin) == (26, 6, 1) and not ls.partition_on_hit_rays  # lidar_gp_2d.hpp:31-40
    assert ls.mapping.type == pygp.MappingType.kInverseSqrt
    rs = pygp.RangeSensorGaussianProcess3Dd.Setting()
    assert (rs.row_group_size, rs.row_overlap_size, rs.col_group_size, rs.col_overlap_size) == (24, 6, 8, 2)  # range_sensor_gp_3d.hpp:33-41
    assert pygp.NoisyInputGaussianProcessD.Setting().max_num_samples == -1
    for cls in ("train", "test", "reset", "is_trained", "setting"):
        assert hasattr(pygp.VanillaGaussianProcessD, cls) and hasattr(pygp.LidarGaussianProcess2Dd, cls)
    mp = pygp.MappingD.Setting()
    mp.type, mp.scale = pygp.MappingType.kInverseSqrt, 1.0
    m = pygp.MappingD(mp)
    assert m.map(4.0) == 0.5 and m.inv(0.5) == 4.0                                # src/mapping.cpp:125-129


@pytest.mark.gpu
@pytest.mark.parametrize("sfx,dtype", [("D", np.float64), ("F", np.float32)])
def test_pybind_vanilla_and_noisy_match_oracle(pygp, oracle, sfx, dtype):
    from tests.util import TOL, err_mean, err_var

    rng = np.random.default_rng(3)
    n, t, d = 400, 900, 2
    x = rng.uniform(-1, 1, (n, d)).astype(dtype)
    y = np.sin(3 * x).sum(axis=1).astype(dtype)
    var = rng.uniform(0.005, 0.02, n).astype(dtype)
    xt = rng.uniform(-1, 1, (t, d)).astype(dtype)
    cls = getattr(pygp, "VanillaGaussianProcess" + sfx)
    s = cls.Setting()
    s.kernel_type, s.max_num_samples = "erl::covariance::Matern32<double, 2>", n
    s.kernel.scale = 0.3
    gp = cls(s)
    assert gp.test(xt.T) is None                                                  # not trained -> None (nullptr)
    assert gp.train(x.T, y[:, None], var) and gp.is_trained                      # x is (x_dim, n), y is (n, y_dim)
    res = gp.test(xt.T)
    o = oracle.VanillaGp(oracle.MATERN32, 0.3, dtype, max_num_samples=n)
    assert o.train(x, y, var) == 0
    m_ref, v_ref = o.test(xt)
    tol = TOL[np.dtype(dtype)]
    assert res.num_test == t and err_mean(res.get_mean(0, True), m_ref) < tol and err_var(res.get_variance(True), v_ref) < tol
    assert gp.cholesky_k_train.shape == (n, n)
    blob = gp.write()
    gp2 = cls(cls.Setting())
    assert gp2.read(blob) and gp2 == gp and np.array_equal(gp2.test(xt.T).get_mean(0, True), res.get_mean(0, True))
    # noisy-input GP with gradient observations on half of the samples
    ncls = getattr(pygp, "NoisyInputGaussianProcess" + sfx)
    ns = ncls.Setting()
    ns.kernel_type = "erl::covariance::RadialBiasFunction2d"
    ns.kernel.scale = 0.5
    g = ncls(ns)
    grad = (3 * np.cos(3 * x)).astype(dtype)                                      # (n, d): d y / d x_k
    flag = (np.arange(n) % 2 == 0).astype(np.int64)
    assert g.train(x.T, y[:, None], grad.T, flag, np.full(n, 0.01, dtype), np.full(n, 0.01, dtype), np.full(n, 0.02, dtype))
    r = g.test(xt.T, True)
    og = oracle.NoisyInputGp(oracle.RBF, 0.5, False, dtype)
    assert og.train(x, y, grad[:, None, :], 0.01, 0.01, 0.02, flag)
    mean_r, grad_r, var_r, gvar_r, cov_r = og.test(xt, True, True)
    assert err_mean(r.get_mean(0, True), mean_r[:, 0]) < tol
    gr, valid = r.get_gradient(0, True)
    assert valid.all() and err_mean(gr.T, grad_r[:, 0, :]) < tol
    assert err_var(r.get_mean_variance(True), var_r) < tol
    assert np.abs(r.get_gradient_variance(True).T - gvar_r).max() / 12.0 < tol and np.abs(r.get_covariance(True).T - cov_r).max() / 12.0 < tol


@pytest.mark.gpu
def test_pybind_lidar(pygp, oracle):
    from tests.util import TOL, err_mean, err_var

    dtype = np.float32
    cls = pygp.LidarGaussianProcess2Df
    s = cls.Setting()
    s.group_size, s.overlap_size = 64, 18
    s.sensor_frame.angle_min, s.sensor_frame.angle_max, s.sensor_frame.num_rays = -3 * np.pi / 4, 3 * np.pi / 4, 1080
    s.sensor_frame.valid_range_min, s.sensor_frame.valid_range_max = 0.1, 30.0
    s.gp.kernel_type = "erl::covariance::OrnsteinUhlenbeck1d"
    s.gp.kernel.scale = 0.05
    lg = cls(s)
    parts = lg.angle_partitions
    assert len(parts) == 24 == lg.num_gps and parts[0][1] - parts[0][0] == 43
    rng = np.random.default_rng(3)
    n = 1080
    ang = (-3 * np.pi / 4 + (3 * np.pi / 2) * np.arange(n, dtype=dtype) / dtype(n - 1)).astype(dtype)  # the frame's own linspace formula
    ranges = (5 + 2 * np.sin(3 * ang)).astype(dtype)
    ranges[rng.random(n) < 0.02] = 1e3
    assert lg.train(np.eye(2, dtype=dtype), np.zeros(2, dtype=dtype), ranges)
    q = rng.uniform(-3 * np.pi / 4 - 0.05, 3 * np.pi / 4 + 0.05, 5000).astype(dtype)
    res = lg.test(q, True, True)
    ok, mean = res.get_mean(True)
    ok2, variance = res.get_variance(True)
    og = oracle.LidarGp2D(ang, oracle.OU, 0.05, 64, 18, 1, True, 0.01, 10.0, False, 2, 1.0, 0.1, 30.0, dtype)
    hit = np.isfinite(ranges) & (ranges >= 0.1) & (ranges <= 30.0)
    assert og.train(ranges, hit)
    m_ref, v_ref, ok_ref = og.test(q, True, True)
    assert np.array_equal(ok, ok_ref) and np.array_equal(ok, ok2) and np.isnan(mean[~ok]).all()
    assert err_mean(mean[ok], m_ref[ok]) < 1e-4 and err_var(variance[ok], v_ref[ok]) < 1e-4
    occ = lg.compute_occ(np.array([2.0, 0.5], dtype=dtype))
    assert set(occ) == {"success", "dist_pos", "range_pred", "occ"} and occ["success"]
