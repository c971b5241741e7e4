"""GPU parity at the sizes BASELINE.json states (VERDICT r01 "next round" item 1a): C5 (n = 16384 blocked Cholesky + predict
slice, SPGP M = 2048 x 2000 samples), C3 at 480 x 640 with both groupings, an ill-conditioned FP32 partition at n = 192 / 256,
and every Mapping type.  The checker is the oracle (CPU port) or LAPACK through scipy where the port would take minutes."""
import numpy as np
import pytest

from tests.util import TOL, err_mean, err_var

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gp():
    import erl_gaussian_process_b200 as m

    return m


def test_vanilla_c5_full_size(gp):
    """BASELINE config 5 (first half): VanillaGp<double>, Matern32 l = 0.1, N = 16384 (32 block columns of the look-ahead
    Cholesky, src/vanilla_gp.cpp:492-505), predict (:134-150) checked on a 4096-point slice of a 65536-point Test() against
    LAPACK potrf / trsm (scipy) on the same inputs; 1e-10 norm-wise (north_star tolerance in double)."""
    import scipy.linalg as sl

    from oracle import oracle_np

    n, t, t_chk = 16384, 65536, 4096
    rng = np.random.default_rng(1)
    x = rng.uniform(-1, 1, (n, 2))
    y = 2 * np.sin(10 * x[:, 0]) * np.cos(10 * x[:, 1])
    var = np.full(n, 1e-3)
    xt = np.random.default_rng(2).uniform(-1, 1, (t, 2))
    g = gp.VanillaGaussianProcess(gp.VanillaGaussianProcess.Setting("matern32", 0.1, max_num_samples=-1), np.float64)
    assert g.train(x, y, var) and g.info == 0
    res = g.test(xt)
    mean, variance = res.get_mean(0), res.get_variance()
    assert np.isfinite(mean).all() and np.isfinite(variance).all()
    assert variance.min() > -1e-9 and variance.max() <= 1.0 + 1e-12

    k = oracle_np.ktrain(oracle_np.MATERN32, 0.1, x, var)
    c = sl.cholesky(k, lower=True, overwrite_a=False, check_finite=False)
    _, l, a = g.get()
    assert np.abs(np.triu(l, 1)).max() == 0  # strict upper triangle zeroed, as matrixL() (src/vanilla_gp.cpp:499)
    el = np.abs(l - c).max() / np.abs(c).max()
    assert el < 1e-11, el
    del l, k
    alpha = sl.cho_solve((c, True), y, check_finite=False)
    ea = np.abs(a[:, 0] - alpha).max() / np.abs(alpha).max()
    assert ea < 1e-8, ea  # alpha carries cond(K) ~ 1e5
    sel = np.random.default_rng(3).choice(t, t_chk, replace=False)
    kt = oracle_np.ktest(oracle_np.MATERN32, 0.1, x, xt[sel])
    m_ref = kt.T @ alpha
    v = sl.solve_triangular(c, kt, lower=True, check_finite=False, overwrite_b=True)
    v_ref = 1.0 - (v * v).sum(axis=0)
    em, ev = err_mean(mean[sel], m_ref), err_var(variance[sel], v_ref)
    assert em < 1e-10, em
    assert ev < 1e-10, ev
    # size-independent properties on the whole Test(): predicting at training points reproduces y to the noise level and
    # drives the variance down to ~ var / (1 + var)
    res_tr = g.test(x[:8192])
    assert np.abs(res_tr.get_mean(0) - y[:8192]).max() < 0.05
    assert res_tr.get_variance().max() < 2e-3


def test_spgp_m2048(gp, oracle):
    """BASELINE config 5 (second half): SPGP with M = 2048 pseudo-inputs, 2000 samples per update, Matern32 l = 0.18, noise 1e-4
    (config/spgp_occupancy_map_2d.yaml:2-20), 100 x 100 test grid (test_spgp_occupancy_map_2d.cpp:366-367); three incremental
    updates vs the CPU port (src/sparse_pseudo_input_gp.cpp:751-791, 835-842, 133-163, 280-310)."""
    gx, gy = np.linspace(-3, 3, 64), np.linspace(-3, 3, 32)
    z = np.array([[a, b] for a in gx for b in gy])
    assert len(z) == 2048
    rng = np.random.default_rng(7)
    gt = np.linspace(-3, 3, 100)
    xt = np.array([[a, b] for a in gt for b in gt])
    g = gp.SparsePseudoInputGaussianProcess("matern32", 0.18, z, np.float64)
    o = oracle.Spgp(oracle.MATERN32, 0.18, z, np.float64)
    for _ in range(3):
        x = rng.uniform(-3, 3, (2000, 2))
        yv = np.tanh(x[:, 0] * x[:, 1])
        var = np.full(2000, 1e-4)
        assert g.update(x, yv, var) and o.update(x, yv, var)
    mean, variance = g.test(xt)
    m_ref, v_ref = o.test(xt)
    em, ev = err_mean(mean, m_ref), err_var(variance, v_ref)
    assert em < 1e-10, em
    assert ev < 1e-10, ev


def _image(rng, rows, cols, dtype):
    r, c = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
    img = 4.0 + 0.8 * np.sin(r / 9.0) * np.cos(c / 13.0) + 0.002 * c
    img[rng.random((rows, cols)) < 0.05] = np.inf
    return img.astype(dtype)


@pytest.mark.parametrize("rg,ro,cg,co,grid,nmax", [(24, 6, 8, 2, (27, 107), 192), (16, 2, 16, 2, (35, 46), 256), (22, 2, 22, 2, (25, 33), 484)])
def test_range_sensor_3d_c3_full_size(gp, oracle, rg, ro, cg, co, grid, nmax):
    """BASELINE config 3: RangeSensorGp3D<float> on a 480 x 640 range image, Matern32 l = 0.05, predict at every pixel
    direction (T = 307 200).  Both groupings SURVEY.md 8(d) recommends: the reference defaults (24,6) x (8,2) -> 27 x 107 =
    2889 GPs with n <= 192, and (16,2)^2 -> 35 x 46 = 1610 GPs with n = 256 (src/range_sensor_gp_3d.cpp:199-259, 321-407); and
    the nearest reachable grid to BASELINE's literal "32 x 24": step 20 -> 33 x 25 = 825 GPs of n = 22 x 22 = 484 samples, beyond
    the shared-memory kernels (large-GP path, erl_gp_largegp.cu)."""
    dtype = np.float32
    rows, cols = 480, 640
    s = gp.RangeSensorGaussianProcess3D.Setting()
    s.row_group_size, s.row_overlap_size, s.col_group_size, s.col_overlap_size = rg, ro, cg, co
    s.sensor_frame.azimuth_min, s.sensor_frame.azimuth_max, s.sensor_frame.num_azimuth_lines = -0.6, 0.6, rows
    s.sensor_frame.elevation_min, s.sensor_frame.elevation_max, s.sensor_frame.num_elevation_lines = -0.8, 0.8, cols
    s.sensor_frame.valid_range_min, s.sensor_frame.valid_range_max = 0.1, 30.0
    s.gp.kernel_type, s.gp.scale = "matern32", 0.05
    rg3 = gp.RangeSensorGaussianProcess3D(s, dtype)
    assert tuple(rg3.grid) == grid
    fc = rg3.sensor_frame.frame_coords
    og = oracle.RangeSensorGp3D(fc, oracle.KERNELS["matern32"], 0.05, rg, ro, 0, cg, co, 0, 32, 0.01, 2, 1.0, dtype)
    assert tuple(og.grid) == grid
    img = _image(np.random.default_rng(5), rows, cols, dtype)
    assert rg3.train(np.eye(3), np.zeros(3), img)
    frame = rg3.sensor_frame
    assert og.train(frame.ranges, frame.mask_hit)
    nr, nc = rg3.grid
    n_seen = 0
    for gidx in np.random.default_rng(6).choice(nr * nc, 40, replace=False):
        info, n, l, a = rg3.get_gp(int(gidx % nr), int(gidx // nr))
        tr, n_ref, l_ref, _ = og.get_gp(int(gidx))
        assert (info == 0) == tr and n == n_ref and n <= nmax
        if tr:
            n_seen = max(n_seen, n)
            assert np.abs(l - l_ref).max() / np.abs(l_ref).max() < 5e-5
    assert n_seen > nmax * 0.8  # the large instances (NBLK 12 / 16) really ran
    coords = fc.reshape(-1, 2).copy()
    res = rg3.test_frame_coords(coords, None, True)
    mean, valid = res.get_mean()
    variance, _ = res.get_variance()
    m_ref, v_ref, ok_ref = og.test(coords, None, True)
    assert np.array_equal(valid, ok_ref) and valid.sum() > 0.95 * len(coords)
    em, ev = err_mean(mean[valid], m_ref[valid]), err_var(variance[valid], v_ref[valid])
    assert em < 1e-4, em
    assert ev < 1e-4, ev


@pytest.mark.parametrize("pr,pc", [(24, 8), (16, 16)])
def test_batch_f32_ill_conditioned_partition(gp, oracle, pr, pc):
    """cond(K) ~ 1e4 in FP32 (pixel patch with 1.6 mrad pitch, Matern32 l = 0.05, noise 0.01) at the
    partition sizes of C3, n = 192 and n = 256: the 3xTF32 products, sqrt.approx / ex2.approx covariance entries and rsqrt
    pivots of the row-GP kernel must stay inside the 1e-4 budget against the FP32 port AND the FP64 port."""
    n = pr * pc
    b = 96
    rng = np.random.default_rng(40 + n)
    pitch = 1.6e-3
    r, c = np.meshgrid(np.arange(pr), np.arange(pc), indexing="ij")
    base = np.stack([r.ravel() * pitch, c.ravel() * pitch], axis=1)
    x = (base[None] + rng.uniform(-0.3, 0.3, (b, 1, 2))).astype(np.float32)
    rngs = 3.0 + rng.uniform(0, 2, (b, 1)) + 0.3 * np.sin(40 * x[..., 0] + rng.uniform(0, 6, (b, 1))) * np.cos(55 * x[..., 1])
    y = (1.0 / np.sqrt(rngs)).astype(np.float32)
    var = np.full((b, n), 0.01, dtype=np.float32)
    n_train = np.full(b, n, dtype=np.int32)
    t = 160
    q_offsets = np.arange(b + 1, dtype=np.int64) * t
    lo, hi = x.min(axis=1, keepdims=True), x.max(axis=1, keepdims=True)
    q_x = (lo + (hi - lo) * rng.random((b, t, 2))).astype(np.float32).reshape(-1, 2)
    from oracle import oracle_np

    k64 = oracle_np.ktrain(oracle_np.MATERN32, 0.05, x[0].astype(np.float64), var[0].astype(np.float64))
    assert np.linalg.cond(k64) > 3e3  # the point of the test
    out = gp.BatchGp(b, n, 2, "matern32", 0.05, np.float32).train_predict(n_train, x, y, var, q_offsets, q_x)
    assert (out["info"] == 0).all() and out["valid"].all()
    ref32 = oracle.batched_train_predict(oracle.MATERN32, 0.05, n_train, x, y, var, q_offsets, q_x)
    ref64 = oracle.batched_train_predict(oracle.MATERN32, 0.05, n_train, x.astype(np.float64), y.astype(np.float64), var.astype(np.float64), q_offsets, q_x.astype(np.float64))
    for ref in (ref32, ref64):
        em, ev = err_mean(out["mean"], ref["mean"]), err_var(out["var"], ref["var"])
        assert em < 1e-4, em
        assert ev < 1e-4, ev


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("mapping,mscale", [(0, 1.0), (1, 1.0), (2, 1.0), (3, 0.2), (4, 0.5), (5, 0.1), (6, 0.3)])
def test_lidar_all_mapping_types(gp, oracle, dtype, mapping, mscale):
    """Mapping<Dtype>::map before training and ::inv after predict for every type of src/mapping.cpp:112-164
    (kIdentity, kInverse, kInverseSqrt, kExp, kLog, kTanh, kSigmoid) through LidarGaussianProcess2D::Train / Test."""
    rng = np.random.default_rng(90 + mapping)
    n = 360
    ang = np.linspace(-2.0, 2.0, n).astype(dtype)
    ranges = (5 + 2 * np.sin(3 * ang) + 0.3 * np.cos(11 * ang)).astype(dtype)
    ranges[rng.random(n) < 0.03] = 1e3
    s = gp.LidarGaussianProcess2D.Setting()
    s.group_size, s.overlap_size, s.margin, s.symmetric_partitions = 40, 10, 1, True
    s.sensor_range_var = 0.01
    s.sensor_frame.angle_min, s.sensor_frame.angle_max, s.sensor_frame.num_rays = float(ang[0]), float(ang[-1]), n
    s.sensor_frame.valid_range_min, s.sensor_frame.valid_range_max = 0.1, 30.0
    s.gp.kernel_type, s.gp.scale = "matern32", 0.1
    s.mapping_type, s.mapping_scale = mapping, mscale
    lg = gp.LidarGaussianProcess2D(s, dtype)
    lg.sensor_frame.angles = ang
    og = oracle.LidarGp2D(ang, oracle.KERNELS["matern32"], 0.1, 40, 10, 1, True, 0.01, 10.0, False, mapping, mscale, 0.1, 30.0, dtype)
    assert lg.train(np.eye(2), np.zeros(2), ranges)
    frame = lg.sensor_frame
    assert og.train(frame.ranges, frame.mask_hit, frame.mask_continuous)
    q = rng.uniform(ang[2], ang[-3], 4000).astype(dtype)
    tol = TOL[np.dtype(dtype)]
    for un_map in (False, True):
        res = lg.test(q, True, un_map)
        mean, valid = res.get_mean()
        variance, _ = res.get_variance()
        m_ref, v_ref, ok_ref = og.test(q, True, un_map)
        assert np.array_equal(valid, ok_ref) and valid.sum() > 3000
        # inv() of a steep map amplifies the mapped-mean error by |inv'(f)|: the gate on the un-mapped mean is the north_star
        # tolerance times that measured amplification (1 for un_map = False and for the identity)
        amp = 1.0
        if un_map and mapping != 0:
            f = og.test(q, True, False)[0][ok_ref].astype(np.float64)
            r = m_ref[ok_ref].astype(np.float64)
            d_inv = {1: lambda v: 1 / v ** 2, 2: lambda v: 2 / np.abs(v) ** 3, 3: lambda v: 1 / (mscale * np.abs(v)), 4: lambda v: np.exp(v) / mscale,
                     5: lambda v: 1 / (mscale * (1 - v ** 2)), 6: lambda v: 1 / (mscale * v * (1 - v))}[mapping](f)
            amp = max(1.0, float((d_inv * np.abs(f).max()).max() / np.abs(r).max()))
        assert err_mean(mean[valid], m_ref[valid]) < tol * amp, (mapping, un_map, amp)
        assert err_var(variance[valid], v_ref[valid]) < tol
        assert np.isfinite(mean[valid]).all()
