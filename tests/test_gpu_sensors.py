"""GPU parity: LidarGaussianProcess2D / RangeSensorGaussianProcess3D (partitioned small GPs) vs the oracle."""
import os

import numpy as np
import pytest

from tests.util import TOL, err_mean, err_var

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lidar_train_frames.npz")


@pytest.fixture(scope="module")
def gp():
    import erl_gaussian_process_b200 as m

    return m


def _c2_scan(rng, n=1080, dtype=np.float32):
    """1080 beams over [-3pi/4, 3pi/4], smooth range profile, ~2 % misses."""
    ang = np.linspace(-3 * np.pi / 4, 3 * np.pi / 4, n)
    rng_ = 5 + 2 * np.sin(3 * ang) + 0.5 * np.sign(np.sin(7 * ang))
    miss = rng.random(n) < 0.02
    rng_[miss] = 1e3  # out of range
    return ang.astype(dtype), rng_.astype(dtype)


def _make_lidar(gp, oracle, dtype, angles, group_size, overlap_size, symmetric, kernel, scale, mapping, discon=False, range_max=30.0, on_hit_rays=False):
    s = gp.LidarGaussianProcess2D.Setting()
    s.partition_on_hit_rays = on_hit_rays
    s.group_size, s.overlap_size, s.margin, s.symmetric_partitions = group_size, overlap_size, 1, symmetric
    s.sensor_range_var, s.discontinuity_var = 0.01, 100.0
    s.sensor_frame.angle_min, s.sensor_frame.angle_max, s.sensor_frame.num_rays = float(angles[0]), float(angles[-1]), len(angles)
    s.sensor_frame.valid_range_min, s.sensor_frame.valid_range_max = 0.1, range_max
    s.sensor_frame.discontinuity_detection = discon
    s.gp.kernel_type, s.gp.scale = kernel, scale
    s.mapping_type = mapping
    s.sensor_frame.angles = np.asarray(angles, dtype=dtype)  # the test's / the log's own angles, not the stand-in frame's linspace
    lg = gp.LidarGaussianProcess2D(s, dtype)
    og = oracle.LidarGp2D(lg.sensor_frame.angles, oracle.KERNELS[kernel], scale, group_size, overlap_size, 1, symmetric, 0.01, 100.0, discon, mapping, 1.0, 0.1, 30.0, dtype,
                          partition_on_hit_rays=on_hit_rays)
    return lg, og


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_lidar_c2_synthetic(gp, oracle, dtype):
    rng = np.random.default_rng(3)
    ang, ranges = _c2_scan(rng, dtype=dtype)
    lg, og = _make_lidar(gp, oracle, dtype, ang, 64, 18, True, "ou", 0.05, 2)
    assert lg.num_partitions == 24 == og.num_partitions
    sizes = [b - a for a, b, _, _ in lg.angle_partitions]
    assert sizes[0] == 43 and sizes[-1] == 43 and all(v == 64 for v in sizes[1:-1])
    assert lg.train(np.eye(2), np.zeros(2), ranges)
    frame = lg.sensor_frame
    assert og.train(frame.ranges, frame.mask_hit, frame.mask_continuous)
    ltol = 2e-5 if dtype == np.float32 else 1e-11
    for p in range(24):
        info, n, l, a = lg.get_gp(p)
        tr, n_ref, l_ref, a_ref = og.get_gp(p)
        assert (info == 0) == tr and n == n_ref
        assert np.abs(l - l_ref).max() / np.abs(l_ref).max() < ltol
    # 100k local test rays (BASELINE config 2) incl. some outside the field of view and a NaN
    t = 100_000
    q = rng.uniform(-3 * np.pi / 4 - 0.05, 3 * np.pi / 4 + 0.05, t).astype(dtype)
    q[17] = np.nan
    res = lg.test(q, True, True)
    mean, valid = res.get_mean()
    var, valid2 = res.get_variance()
    m_ref, v_ref, ok_ref = og.test(q, True, True)
    assert np.array_equal(valid, ok_ref) and np.array_equal(valid, valid2)
    assert not valid[17] and (~valid).sum() > 100
    assert np.isnan(mean[~valid]).all()  # untouched
    tol = TOL[np.dtype(dtype)]
    assert err_mean(mean[valid], m_ref[valid]) < tol
    assert err_var(var[valid], v_ref[valid]) < tol
    # un_map = False returns the mapped quantity (1/sqrt(range))
    m2, _ = lg.test(q[:5000], True, False).get_mean()
    m2_ref, _, ok2 = og.test(q[:5000], True, False)
    assert err_mean(m2[ok2], m2_ref[ok2]) < tol


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_lidar_reference_log_frames(gp, oracle, dtype):
    """Frames of the reference's own lidar log, configured as its gtest (test_lidar_gp_2d.cpp:107-158):
    group 26 / overlap 6, asymmetric partitions, OU l = 0.05, identity mapping, predict at the training angles."""
    data = np.load(GOLDEN)
    for k in range(len(data["frame_ids"])):
        ang, ranges = data["angles"][k].astype(dtype), data["ranges"][k].astype(dtype)
        lg, og = _make_lidar(gp, oracle, dtype, ang, 26, 6, False, "ou", 0.05, 0)
        assert lg.num_partitions == og.num_partitions
        assert lg.train(np.eye(2), np.zeros(2), ranges)
        frame = lg.sensor_frame
        assert og.train(frame.ranges, frame.mask_hit, frame.mask_continuous)
        res = lg.test(ang, True, True)
        mean, valid = res.get_mean()
        var, _ = res.get_variance()
        m_ref, v_ref, ok_ref = og.test(ang, True, True)
        assert valid.any() and np.array_equal(valid, ok_ref)  # the reference asserts success.any()
        tol = TOL[np.dtype(dtype)]
        assert err_mean(mean[valid], m_ref[valid]) < tol
        assert err_var(var[valid], v_ref[valid]) < tol
        mae = np.abs(mean[valid] - ranges[valid]).mean()
        assert mae < 0.08  # the reference's own threshold, test_lidar_gp_2d.cpp:261


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_lidar_partition_on_hit_rays(gp, oracle, dtype):
    """Setting::partition_on_hit_rays (src/lidar_gp_2d.cpp:302-348, 364): the table follows the hit rays of every frame.  Three
    frames through ONE object (few misses / 30 % misses / last rays missing - the regime where the reference's own index
    arithmetic stays in bounds), tables and partition GPs against the oracle, then the reference log's frames."""
    rng = np.random.default_rng(21)
    ang, ranges0 = _c2_scan(rng, n=540, dtype=dtype)
    lg, og = _make_lidar(gp, oracle, dtype, ang, 40, 10, True, "ou", 0.05, 2, on_hit_rays=True)
    assert lg.num_partitions == 0 == og.num_partitions
    tol = TOL[np.dtype(dtype)]
    for frame_id, miss in enumerate([0.02, 0.3, 0.1]):
        ranges = ranges0.copy()
        ranges[rng.random(len(ang)) < miss] = 1e3
        if frame_id == 2:
            ranges[-4:] = 1e3
        assert lg.train(np.eye(2), np.zeros(2), ranges)
        frame = lg.sensor_frame
        assert og.train(frame.ranges, frame.mask_hit, frame.mask_continuous)
        parts, parts_ref = lg.angle_partitions, og.angle_partitions
        assert len(parts) == len(parts_ref) > 3
        assert [(a, b) for a, b, _, _ in parts] == [(a, b) for a, b, _, _ in parts_ref]
        assert all(c1 == c2 and d1 == d2 for (_, _, c1, d1), (_, _, c2, d2) in zip(parts, parts_ref))
        for p in range(len(parts)):
            info, n, l, a = lg.get_gp(p)
            tr, n_ref, l_ref, a_ref = og.get_gp(p)
            assert (info == 0) == tr and n == n_ref
            if tr:
                assert np.abs(l - l_ref).max() / np.abs(l_ref).max() < (2e-5 if dtype == np.float32 else 1e-11)
        q = rng.uniform(ang[0] - 0.05, ang[-1] + 0.05, 20000).astype(dtype)
        res = lg.test(q, True, True)
        mean, valid = res.get_mean()
        var, _ = res.get_variance()
        m_ref, v_ref, ok_ref = og.test(q, True, True)
        assert np.array_equal(valid, ok_ref) and valid.sum() > 10000
        assert err_mean(mean[valid], m_ref[valid]) < tol
        assert err_var(var[valid], v_ref[valid]) < tol
    data = np.load(GOLDEN)
    for k in range(min(4, len(data["frame_ids"]))):
        a, r = data["angles"][k].astype(dtype), data["ranges"][k].astype(dtype)
        lg, og = _make_lidar(gp, oracle, dtype, a, 26, 6, False, "ou", 0.05, 0, on_hit_rays=True)
        assert lg.train(np.eye(2), np.zeros(2), r)
        assert og.train(lg.sensor_frame.ranges, lg.sensor_frame.mask_hit, lg.sensor_frame.mask_continuous)
        assert [(x, y) for x, y, _, _ in lg.angle_partitions] == [(x, y) for x, y, _, _ in og.angle_partitions]
        mean, valid = lg.test(a, True, True).get_mean()
        m_ref, _, ok_ref = og.test(a, True, True)
        assert valid.any() and np.array_equal(valid, ok_ref)
        assert err_mean(mean[valid], m_ref[valid]) < tol


def test_lidar_world_frame_angles_and_discontinuity(gp, oracle):
    dtype = np.float64
    rng = np.random.default_rng(5)
    ang, ranges = _c2_scan(rng, n=540, dtype=dtype)
    lg, og = _make_lidar(gp, oracle, dtype, ang, 40, 10, True, "matern32", 0.1, 2, discon=True)
    th = 0.7
    rot = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    assert lg.train(rot, np.array([1.0, 2.0]), ranges)
    frame = lg.sensor_frame
    assert (~frame.mask_continuous).any()
    assert og.train(frame.ranges, frame.mask_hit, frame.mask_continuous, rotation=rot)
    q_local = rng.uniform(ang[2], ang[-3], 5000)
    q_world = q_local + th
    mean, valid = lg.test(q_world, False, True).get_mean()
    var, _ = lg.test(q_world, False, True).get_variance()
    m_ref, v_ref, ok_ref = og.test(q_world, False, True)
    # world->frame goes through cos/sin/atan2 on both sides: rays within 1e-9 rad of a partition edge may flip
    edges = np.array([c for _, _, cl, cr in lg.angle_partitions for c in (cl, cr)])
    safe = np.abs(q_local[:, None] - edges[None, :]).min(axis=1) > 1e-9
    assert np.array_equal(valid[safe], ok_ref[safe])
    both = valid & ok_ref & safe
    assert err_mean(mean[both], m_ref[both]) < 1e-9
    assert err_var(var[both], v_ref[both]) < 1e-9


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_lidar_compute_occ(gp, oracle, dtype):
    rng = np.random.default_rng(8)
    ang, ranges = _c2_scan(rng, n=360, dtype=dtype)
    lg, og = _make_lidar(gp, oracle, dtype, ang, 26, 6, True, "ou", 0.05, 2)
    assert lg.train(np.eye(2), np.zeros(2), ranges)
    frame = lg.sensor_frame
    og.train(frame.ranges, frame.mask_hit, frame.mask_continuous)
    a = rng.uniform(-3.0, 3.0, 400)
    d = rng.uniform(0.5, 8.0, 400)
    pos = np.stack([d * np.cos(a), d * np.sin(a)], axis=1).astype(dtype)
    ok, dist, rp, occ = lg.compute_occ(pos)
    n_ok = 0
    for i in range(len(pos)):
        r_ok, r_d, r_rp, r_occ = og.compute_occ(float(pos[i, 0]), float(pos[i, 1]))
        edge = min(abs(np.arctan2(pos[i, 1], pos[i, 0]) - c) for _, _, cl, cr in lg.angle_partitions for c in (cl, cr)) < 1e-5
        if edge:
            continue
        assert ok[i] == r_ok, i
        if r_ok:
            n_ok += 1
            tol = 2e-4 if dtype == np.float32 else 1e-9
            assert abs(dist[i] - r_d) < tol * max(1, r_d)
            assert abs(rp[i] - r_rp) < tol * max(1, abs(r_rp))
            # |d occ / d f| <= a / 2 with a = 30 d (src/lidar_gp_2d.cpp:455-457): the mean budget (1e-4 / 1e-10, relative to f ~ 1 / sqrt(range)
            # <= 1.5 here) times that slope bounds the occ difference of two correct implementations
            assert abs(occ[i] - r_occ) <= 0.5 * 30.0 * r_d * TOL[np.dtype(dtype)] * 1.5 + (1e-6 if dtype == np.float32 else 1e-12)
    assert n_ok > 50 and (~ok).sum() > 10


def _range_image(rng, rows, cols, dtype):
    """plane + sinusoidal relief + 5 % invalid pixels."""
    r, c = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
    img = 4.0 + 0.8 * np.sin(r / 9.0) * np.cos(c / 13.0) + 0.002 * c
    img[rng.random((rows, cols)) < 0.05] = np.inf
    return img.astype(dtype)


@pytest.mark.parametrize("dtype,kernel,rg,ro,cg,co", [(np.float32, "matern32", 24, 6, 8, 2), (np.float64, "ou", 12, 2, 10, 4), (np.float32, "matern32", 16, 2, 16, 2)])
def test_range_sensor_3d(gp, oracle, dtype, kernel, rg, ro, cg, co):
    rng = np.random.default_rng(15)
    rows, cols = 96, 128
    s = gp.RangeSensorGaussianProcess3D.Setting()
    s.row_group_size, s.row_overlap_size, s.col_group_size, s.col_overlap_size = rg, ro, cg, co
    s.sensor_frame.azimuth_min, s.sensor_frame.azimuth_max, s.sensor_frame.num_azimuth_lines = -0.6, 0.6, rows
    s.sensor_frame.elevation_min, s.sensor_frame.elevation_max, s.sensor_frame.num_elevation_lines = -0.4, 0.4, cols
    s.sensor_frame.valid_range_min, s.sensor_frame.valid_range_max = 0.1, 30.0
    s.gp.kernel_type, s.gp.scale = kernel, 0.05
    rg3 = gp.RangeSensorGaussianProcess3D(s, dtype)
    fc = rg3.sensor_frame.frame_coords
    og = oracle.RangeSensorGp3D(fc, oracle.KERNELS[kernel], 0.05, rg, ro, 0, cg, co, 0, 32, 0.01, 2, 1.0, dtype)
    assert rg3.grid == og.grid
    for axis in (0, 1):
        parts = rg3.partitions(axis)
        coords = fc[:, 0, 0] if axis == 0 else fc[0, :, 1]
        ref = oracle.make_partitions(coords, (rg, cg)[axis], (ro, co)[axis], 0, True)
        assert [(a, b) for a, b, _, _ in parts] == [(a, b) for a, b, _, _ in ref]
        assert all(p[2] == q[2] and p[3] == q[3] for p, q in zip(parts, ref))
    img = _range_image(rng, rows, cols, dtype)
    assert rg3.train(np.eye(3), np.zeros(3), img)
    frame = rg3.sensor_frame
    assert og.train(frame.ranges, frame.mask_hit)
    nr, nc = rg3.grid
    ltol = 5e-5 if dtype == np.float32 else 1e-10
    n_trained = 0
    for g in rng.choice(nr * nc, 25, replace=False):
        info, n, l, a = rg3.get_gp(int(g % nr), int(g // nr))
        tr, n_ref, l_ref, a_ref = og.get_gp(int(g))
        assert (info == 0) == tr and n == n_ref
        if tr:
            n_trained += 1
            assert np.abs(l - l_ref).max() / np.abs(l_ref).max() < ltol
    assert n_trained > 0
    # full-image predict: every pixel direction, plus rays outside the frame and flagged-bad rays
    coords = fc.reshape(-1, 2).copy()
    extra = np.array([[5.0, 0.0], [0.0, 5.0], [np.nan, 0.0]], dtype=dtype)
    coords = np.concatenate([coords, extra])
    ok_in = np.ones(len(coords), dtype=bool)
    ok_in[11] = False
    res = rg3.test_frame_coords(coords, ok_in, True)
    mean, valid = res.get_mean()
    var, _ = res.get_variance()
    m_ref, v_ref, ok_ref = og.test(coords, ok_in, True)
    assert np.array_equal(valid, ok_ref)
    assert not valid[11] and not valid[-3:].any()
    tol = TOL[np.dtype(dtype)]
    assert err_mean(mean[valid], m_ref[valid]) < tol
    assert err_var(var[valid], v_ref[valid]) < tol
    # Test(directions): direction vectors through the frame stand-in
    az, el = coords[:200, 0].astype(np.float64), coords[:200, 1].astype(np.float64)
    dirs = np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)], axis=1)
    m3, v3 = rg3.test(dirs, True, True).get_mean()
    assert v3.sum() > 150


def test_range_sensor_3d_limits(gp):
    s = gp.RangeSensorGaussianProcess3D.Setting()
    s.row_overlap_size = 3  # odd overlap: the reference constructor asserts (src/range_sensor_gp_3d.cpp:190-197)
    with pytest.raises(ValueError):
        gp.RangeSensorGaussianProcess3D(s, np.float32)
    s = gp.RangeSensorGaussianProcess3D.Setting()
    s.row_group_size, s.col_group_size = 48, 44  # 2112 samples per GP: beyond the large-GP path (2048)
    with pytest.raises(gp.ErlGpError):
        gp.RangeSensorGaussianProcess3D(s, np.float32)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_range_sensor_3d_compute_occ(gp, oracle, dtype):
    """Batched 3-D ComputeOcc (src/range_sensor_gp_3d.cpp:409-439): variance gate, mapped mean, occ, un-mapped range."""
    rng = np.random.default_rng(21)
    rows, cols = 64, 96
    s = gp.RangeSensorGaussianProcess3D.Setting()
    s.row_group_size, s.row_overlap_size, s.col_group_size, s.col_overlap_size = 12, 2, 10, 4
    s.sensor_frame.azimuth_min, s.sensor_frame.azimuth_max, s.sensor_frame.num_azimuth_lines = -0.5, 0.5, rows
    s.sensor_frame.elevation_min, s.sensor_frame.elevation_max, s.sensor_frame.num_elevation_lines = -0.3, 0.3, cols
    s.sensor_frame.valid_range_min, s.sensor_frame.valid_range_max = 0.1, 30.0
    s.gp.kernel_type, s.gp.scale = "matern32", 0.05
    s.max_valid_range_var = 0.02
    rg3 = gp.RangeSensorGaussianProcess3D(s, dtype)
    fc = rg3.sensor_frame.frame_coords
    og = oracle.RangeSensorGp3D(fc, oracle.KERNELS["matern32"], 0.05, 12, 2, 0, 10, 4, 0, 32, 0.01, 2, 1.0, dtype)
    img = _range_image(rng, rows, cols, dtype)
    assert rg3.train(np.eye(3), np.zeros(3), img)
    assert og.train(rg3.sensor_frame.ranges, rg3.sensor_frame.mask_hit)
    t = 3000
    az = rng.uniform(-0.6, 0.6, t)
    el = rng.uniform(-0.36, 0.36, t)
    d = rng.uniform(0.5, 8.0, t)
    pos = np.stack([d * np.cos(el) * np.cos(az), d * np.cos(el) * np.sin(az), d * np.sin(el)], axis=1).astype(dtype)
    pos[5] = 0  # zero vector: ComputeFrameCoords fails
    ok, dist, rp, occ = rg3.compute_occ(pos)
    ok_c, dist_ref, coords = rg3.sensor_frame.compute_frame_coords(pos)
    m_ref, v_ref, valid_ref = og.test(coords, ok_c, False)  # mapped mean + variance
    with np.errstate(invalid="ignore", over="ignore"):
        good_ref = valid_ref & ~(v_ref > s.max_valid_range_var)
        occ_ref = 2.0 / (1.0 + np.exp(dist_ref.astype(np.float64) * s.occ_test_temperature * (m_ref.astype(np.float64) - 1.0 / np.sqrt(dist_ref.astype(np.float64))))) - 1.0
        rp_ref = 1.0 / (m_ref.astype(np.float64) ** 2)
    # positions whose variance sits within rounding of the gate may flip
    near_gate = valid_ref & (np.abs(v_ref - s.max_valid_range_var) < (1e-5 if dtype == np.float32 else 1e-11))
    sel = ~near_gate
    assert np.array_equal(ok[sel], good_ref[sel])
    assert not ok[5] and ok.sum() > 300 and (~ok).sum() > 100
    both = ok & good_ref
    tol = 2e-4 if dtype == np.float32 else 1e-9
    assert np.abs(rp[both] - rp_ref[both]).max() / np.abs(rp_ref[both]).max() < tol
    # occ = 2 / (1 + exp(a (f - map(d)))) - 1 with a = 30 d: |d occ / d f| <= a / 2, so a mean inside the north_star budget
    # (1e-4 / 1e-10 of max|f|, the same budget test_range_sensor_3d holds the mean to) moves occ by at most (a / 2) x that budget.
    # The bound is per position (the measured worst case is printed).
    budget = TOL[np.dtype(dtype)] * np.abs(m_ref[both]).max()
    lipschitz = 0.5 * dist_ref[both].astype(np.float64) * s.occ_test_temperature
    print(f"ComputeOcc 3-D {np.dtype(dtype).name}: max |occ - occ_ref| = {np.abs(occ[both] - occ_ref[both]).max():.2e}, largest slope a / 2 = {lipschitz.max():.0f}")
    assert (np.abs(occ[both] - occ_ref[both]) <= lipschitz * budget + (1e-6 if dtype == np.float32 else 1e-12)).all()
    assert np.isnan(rp[~ok]).all() and np.isnan(occ[~ok]).all()  # untouched
