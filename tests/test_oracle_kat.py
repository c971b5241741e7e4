"""Pin the oracle (C++ restatement and its numpy twin) to the reference's own known-answer values.

KATs are the author's observed values left as comments next to the gtest thresholds:
  test/gtest/test_vanilla_gp.cpp:103   0.00024246430481069056  (1-D RBF, n=100, T=200)
  test/gtest/test_vanilla_gp.cpp:214   0.0005035569336460338   (2-D RBF, 50^2 train, 100^2 test)
  test/gtest/test_vanilla_gp.cpp:366-7 0.0005035569336460478 / 0.0011257545588707807 (2 outputs)
  test/gtest/test_sparse_pseudo_input_gp.cpp:109  0.00013951539277877418 (SPGP 1-D, M=20, N=1000)
"""
import numpy as np
import pytest

from oracle import oracle_np as onp

KAT_SISO = 0.00024246430481069056
KAT_MISO = 0.0005035569336460338
KAT_MIMO = (0.0005035569336460478, 0.0011257545588707807)
KAT_SPGP = 0.00013951539277877418
NOISE = 0.001


def _grid(nx, ny):
    gx = np.linspace(-1, 1, nx)
    gy = np.linspace(-1, 1, ny)
    pts = np.array([[a, b] for a in gx for b in gy])  # x outer, y inner (test_vanilla_gp.cpp:116-123)
    return pts


def test_vanilla_siso_kat(oracle):
    n, t = 100, 200
    x = np.linspace(0, 2 * np.pi, n)[:, None]
    y = np.sin(x[:, 0])
    xt = np.linspace(0, 2 * np.pi, t)[:, None]
    var = np.full(n, NOISE)
    gp = oracle.VanillaGp(oracle.RBF, 0.5, np.float64, max_num_samples=n)
    assert gp.train(x, y, var) == 0
    mean, variance = gp.test(xt)
    mae = np.abs(mean - np.sin(xt[:, 0])).mean()
    assert mae == pytest.approx(KAT_SISO, rel=1e-9)
    assert mae < 3.0e-4  # the reference's own assertion
    # numpy twin agrees with both
    l, alpha = onp.vanilla_train(onp.RBF, 0.5, x, y, var)
    m2, v2 = onp.vanilla_test(onp.RBF, 0.5, x, l, alpha, xt)
    assert np.abs(m2 - np.sin(xt[:, 0])).mean() == pytest.approx(KAT_SISO, rel=1e-9)
    assert np.abs(mean - m2).max() < 1e-10
    assert np.abs(variance - v2).max() < 1e-10
    assert (variance > -1e-9).all() and (variance < 1).all()


def test_vanilla_miso_mimo_kat(oracle):
    tr = _grid(50, 50)
    te = _grid(100, 100)
    f1 = lambda p: 2 * np.sin(10 * p[:, 0]) * np.cos(10 * p[:, 1])
    f2 = lambda p: 3 * (np.sin(10 * p[:, 0]) + np.cos(10 * p[:, 1]))
    var = np.full(len(tr), NOISE)
    gp = oracle.VanillaGp(oracle.RBF, 0.1, np.float64, max_num_samples=len(tr))
    assert gp.train(tr, np.stack([f1(tr), f2(tr)], axis=1), var) == 0
    mean, _ = gp.test(te, want_var=False)
    mae1 = np.abs(mean[:, 0] - f1(te)).mean()
    mae2 = np.abs(mean[:, 1] - f2(te)).mean()
    assert mae1 == pytest.approx(KAT_MISO, rel=1e-8)
    assert mae1 == pytest.approx(KAT_MIMO[0], rel=1e-8)
    assert mae2 == pytest.approx(KAT_MIMO[1], rel=1e-8)
    assert mae1 < 5.1e-4 and mae2 < 1.2e-3  # reference thresholds :213-215, :363-367


def test_spgp_siso_kat(oracle):
    m, n, t = 20, 1000, 200
    z = np.linspace(0, 2 * np.pi, m)[:, None]
    x = np.linspace(0, 2 * np.pi, n)[:, None]
    y = np.sin(x[:, 0])
    xt = np.linspace(0, 2 * np.pi, t)[:, None]
    var = np.full(n, NOISE)
    gp = oracle.Spgp(oracle.RBF, 0.6, z, np.float64)
    assert gp.update(x, y, var)
    mean, variance = gp.test(xt)
    mae = np.abs(mean - np.sin(xt[:, 0])).mean()
    # cond(K_M) ~ 1e6: two correct implementations agree to ~5 digits only
    assert mae == pytest.approx(KAT_SPGP, rel=2e-4)
    assert mae < 4.02e-4  # reference threshold :107-111
    m2, v2 = onp.spgp_fit_predict(onp.RBF, 0.6, z, x, y, var, xt)
    assert np.abs(mean - m2).max() < 1e-6
    assert np.abs(variance - v2).max() < 1e-6


@pytest.mark.parametrize("kernel", ["ou", "matern32", "rbf"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_cpp_oracle_matches_numpy_twin(oracle, kernel, dtype):
    rng = np.random.default_rng(7)
    n, t, d = 96, 50, 2
    x = rng.uniform(-1, 1, (n, d)).astype(dtype)
    y = np.sin(3 * x[:, 0]) * np.cos(2 * x[:, 1])
    var = np.full(n, 0.01, dtype=dtype)
    xt = rng.uniform(-1, 1, (t, d)).astype(dtype)
    kid = oracle.KERNELS[kernel]
    scale = 0.4
    k = oracle.gram_train(kid, scale, x, var)
    k2 = onp.ktrain(kid, scale, x, var)
    tol = 2e-6 if dtype == np.float32 else 1e-14
    assert np.abs(k - k2).max() < tol
    assert np.abs(oracle.gram_test(kid, scale, x, xt) - onp.ktest(kid, scale, x, xt)).max() < tol
    gp = oracle.VanillaGp(kid, scale, dtype, max_num_samples=n)
    assert gp.train(x, y.astype(dtype), var) == 0
    mean, variance = gp.test(xt)
    l, alpha = onp.vanilla_train(kid, scale, x.astype(np.float64), y.astype(np.float64), var.astype(np.float64))
    m2, v2 = onp.vanilla_test(kid, scale, x.astype(np.float64), l, alpha, xt.astype(np.float64))
    tol = 1e-4 if dtype == np.float32 else 1e-10
    assert np.abs(mean - m2).max() / max(np.abs(m2).max(), 1e-30) < tol
    assert np.abs(variance - v2).max() < tol


def test_partitions_match_survey_c2(oracle):
    # 1080 beams, group 64 / overlap 18 -> 24 partitions of sizes 43, 22 x 64, 43
    ang = np.linspace(-3 * np.pi / 4, 3 * np.pi / 4, 1080)
    parts = oracle.make_partitions(ang, 64, 18, 1, True)
    sizes = [b - a for a, b, _, _ in parts]
    assert len(parts) == 24 and sizes[0] == 43 and sizes[-1] == 43 and all(s == 64 for s in sizes[1:-1])
    ref = onp.make_partitions(ang, 64, 18, 1, True)
    assert [(a, b) for a, b, _, _ in parts] == [(a, b) for a, b, _, _ in ref]
    for (_, _, cl, cr), (_, _, cl2, cr2) in zip(parts, ref):
        assert cl == cl2 and cr == cr2
    # asymmetric (the reference's lidar test uses it, test_lidar_gp_2d.cpp:156)
    pa = oracle.make_partitions(ang[:270], 26, 6, 1, False)
    ra = onp.make_partitions(ang[:270], 26, 6, 1, False)
    assert [(a, b) for a, b, _, _ in pa] == [(a, b) for a, b, _, _ in ra]


def test_hit_ray_partitions(oracle):
    """PartitionOnHitRays (src/lidar_gp_2d.cpp:302-348): the C++ oracle against the literal numpy restatement.  Where the
    reference's own reads stay in bounds (last ray a miss, few misses) the clamped table IS the reference's; the clamps only
    act where the reference is undefined (IndexError in the strict twin)."""
    rng = np.random.default_rng(11)
    ang = np.linspace(-2.3, 2.3, 270)
    strict_ok = 0
    for trial in range(40):
        hit = rng.random(270) > rng.choice([0.0, 0.02, 0.3])
        if trial % 2 == 0:
            hit[-3:] = False  # the reference reads angles[last hit + 1]: in bounds only if the last ray is a miss
        lg = oracle.LidarGp2D(ang, oracle.OU, 0.05, 26, 6, 1, False, dtype=np.float64, partition_on_hit_rays=True)
        assert lg.num_partitions == 0  # the constructor does not partition (:182)
        ranges = np.where(hit, 5.0 + np.sin(3 * ang), 1e3)
        assert lg.train(ranges, hit)
        ref = onp.make_hit_ray_partitions(ang, hit, 26, 6)
        got = lg.angle_partitions
        assert [(a, b) for a, b, _, _ in got] == [(a, b) for a, b, _, _ in ref]
        assert all(c1 == c2 and d1 == d2 for (_, _, c1, d1), (_, _, c2, d2) in zip(got, ref))
        try:
            strict = onp.strict_hit_ray_partitions(ang, hit, 26, 6)
        except IndexError:
            continue
        strict_ok += 1
        assert [(a, b) for a, b, _, _ in strict] == [(a, b) for a, b, _, _ in ref]
    assert strict_ok >= 5  # the in-bounds regime is exercised


def test_noisy_input_gp_reference_mae_values(oracle):
    """NoisyInputGaussianProcess: the oracle's derivative-augmented Gram (erl_covariance's ComputeKtrainWithGradient is absent)
    reproduces the mean-absolute errors the reference's own gtest printed into its source
    (test/gtest/test_noisy_input_gp.cpp:174-178, 348-349): RBF 1-D, 100 samples of sin(2x) on [0, 2 pi], noise 1e-4 on x, y and
    the gradient, 200 test points.  Agreement to 6+ digits fixes the signs, the row / column layout and the noise model
    K[i][i] = 1 + var_x + var_y, K[g][g] = 1 / l^2 + var_grad."""
    n, t = 100, 200
    x = np.linspace(0, 2 * np.pi, n)[:, None]
    xt = np.linspace(0, 2 * np.pi, t)[:, None]
    y, g = np.sin(2 * x[:, 0]), 2 * np.cos(2 * x[:, 0])
    yt, gt = np.sin(2 * xt[:, 0]), 2 * np.cos(2 * xt[:, 0])
    # with gradient observations (:174-178)
    for scale, mae_ref, mae_grad_ref in [(0.5, 8.523327884661321e-06, 0.0001228380577847092), (0.4, 6.453961399301211e-06, 0.0001125665082426853),
                                         (0.3, 4.4761251597013675e-06, 8.851481085195954e-05), (0.2, 4.1624286843223515e-06, 7.139121709502966e-05),
                                         (0.1, 1.756325489369356e-05, 0.00034785637964318994)]:
        gp = oracle.NoisyInputGp(oracle.RBF, scale, False, np.float64)
        assert gp.train(x, y, g[:, None, None], 1e-4, 1e-4, 1e-4, 1) and gp.m == 2 * n
        mean, grad, _, _, _ = gp.test(xt, True, False)
        mae, mae_grad = np.abs(mean[:, 0] - yt).mean(), np.abs(grad[:, 0, 0] - gt).mean()
        # cond(K) grows with the scale (1e8 at 0.2, 1e11 at 0.5): the printed digits are only reproducible to that accuracy
        rtol = 1e-6 if scale <= 0.2 else 2e-2
        assert abs(mae - mae_ref) / mae_ref < rtol and abs(mae_grad - mae_grad_ref) / mae_grad_ref < rtol, (scale, mae, mae_grad)
    # without gradient observations (:348-349): the gradient is still predicted
    for scale, mae_ref, mae_grad_ref in [(0.5, 0.00019489369361661352, 0.003074427178772044), (0.2, 7.377464439757659e-05, 0.0024347632450979033)]:
        gp = oracle.NoisyInputGp(oracle.RBF, scale, True, np.float64)
        assert gp.train(x, y, None, 1e-4, 1e-4, None, 0) and gp.m == n
        mean, grad, _, _, _ = gp.test(xt, True, False)
        mae, mae_grad = np.abs(mean[:, 0] - yt).mean(), np.abs(grad[:, 0, 0] - gt).mean()
        rtol = 1e-6 if scale <= 0.2 else 2e-2
        assert abs(mae - mae_ref) / mae_ref < rtol and abs(mae_grad - mae_grad_ref) / mae_grad_ref < rtol, (scale, mae, mae_grad)


def test_noisy_input_gp_reference_mae_values_2d(oracle):
    """The 2-D case of the same gtest (:366-410, 552-554): 50 x 50 samples of 2 sin(10 x) cos(5 y) with both partial derivatives
    (7500 x 7500 system), RBF l = 0.1, 100 x 100 test points."""
    def values(n):
        xs, ys = np.linspace(-2, 2, n), np.linspace(-1, 1, n)
        px, py = np.meshgrid(xs, ys, indexing="ij")  # xi outer, yi inner (:361-362)
        pts = np.stack([px.ravel(), py.ravel()], axis=1)
        z = 2 * np.sin(10 * pts[:, 0]) * np.cos(5 * pts[:, 1])
        gx = 20 * np.cos(10 * pts[:, 0]) * np.cos(5 * pts[:, 1])
        gy = -10 * np.sin(10 * pts[:, 0]) * np.sin(5 * pts[:, 1])
        return pts, z, gx, gy

    pts, z, gx, gy = values(50)
    gp = oracle.NoisyInputGp(oracle.RBF, 0.1, False, np.float64)
    assert gp.train(pts, z, np.stack([gx, gy], axis=1)[:, None, :], 1e-4, 1e-4, 1e-4, 1) and gp.m == 7500
    pt, zt, gxt, gyt = values(100)
    mean, grad, _, _, _ = gp.test(pt, True, False)
    mae, mae_x, mae_y = np.abs(mean[:, 0] - zt).mean(), np.abs(grad[:, 0, 0] - gxt).mean(), np.abs(grad[:, 0, 1] - gyt).mean()
    for got, ref in ((mae, 9.516671456234042e-06), (mae_x, 0.00010712550862064423), (mae_y, 0.0002508214688791491)):
        assert abs(got - ref) / ref < 1e-4, (mae, mae_x, mae_y)
