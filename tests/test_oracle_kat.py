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
