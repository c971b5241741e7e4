"""GPU parity: NoisyInputGaussianProcess (GP with noisy inputs and gradient observations, src/noisy_input_gp.cpp) vs the oracle,
whose derivative-augmented Gram matrix is pinned by the MAE values of the reference's own gtest (tests/test_oracle_kat.py)."""
import numpy as np
import pytest

from tests.util import TOL

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gp():
    import erl_gaussian_process_b200 as m

    return m


def _rel(a, b):
    return np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)).max() / max(np.abs(b).max(), 1e-30)


def _problem(rng, n, d, y_dim, frac_grad):
    x = rng.uniform(-1, 1, (n, d))
    w = rng.uniform(1, 3, (y_dim, d))
    y = np.stack([np.sin(x @ w[k]) for k in range(y_dim)], axis=1)                 # (n, y_dim)
    grad = np.stack([np.cos(x @ w[k])[:, None] * w[k][None, :] for k in range(y_dim)], axis=1)  # (n, y_dim, d)
    flag = (rng.random(n) < frac_grad).astype(np.int64)
    return x, y, grad, flag


def _check(gp, oracle, dtype, kernel, scale, n, d, y_dim, frac_grad, t, seed, noise=1e-2):
    rng = np.random.default_rng(seed)
    x, y, grad, flag = _problem(rng, n, d, y_dim, frac_grad)
    vx, vy, vg = rng.uniform(0.5, 1.5, n) * noise, rng.uniform(0.5, 1.5, n) * noise, rng.uniform(0.5, 1.5, n) * noise
    xt = rng.uniform(-1.1, 1.1, (t, d))
    s = gp.NoisyInputGaussianProcess.Setting(kernel, scale)
    g = gp.NoisyInputGaussianProcess(s, dtype)
    assert g.test(xt) is None  # Test() before Train(), src/noisy_input_gp.cpp:905
    assert g.train(x, y, grad, vx, vy, vg, flag)
    o = oracle.NoisyInputGp(oracle.KERNELS[kernel], scale, False, dtype)
    assert o.train(x, y, grad, vx, vy, vg, flag)
    info, k, l, a = g.get()
    info_r, k_r, l_r, a_r = o.get()
    m = n + d * int(flag.sum())
    assert info == 0 == info_r and k.shape == (m, m) == k_r.shape
    f32 = np.dtype(dtype) == np.float32
    assert _rel(k, k_r) < (1e-6 if f32 else 1e-14)
    assert np.abs(np.triu(l, 1)).max() == 0
    assert _rel(l, l_r) < (5e-5 if f32 else 1e-11)
    assert _rel(a, a_r) < (5e-3 if f32 else 1e-8)  # alpha carries cond(K)
    res = g.test(xt, True)
    mean_r, grad_r, var_r, gvar_r, cov_r = o.test(xt, True, True)
    tol = TOL[np.dtype(dtype)]
    prior = 3.0 / scale**2  # the scale of the gradient variances / covariances (m_three_over_scale_square_)
    for c in range(y_dim):
        assert _rel(res.get_mean(c), mean_r[:, c]) < tol
        gr, valid = res.get_gradient(c)
        assert valid.all() and _rel(gr, grad_r[:, c, :]) < tol
    assert np.abs(res.get_mean_variance() - var_r).max() < tol
    assert np.abs(res.get_gradient_variance() - gvar_r).max() / prior < tol
    assert np.abs(res.get_covariance() - cov_r).max() / prior < tol
    # value-only TestResult (will_predict_gradient = false)
    res0 = g.test(xt[:100], False)
    m0, _, v0, _, _ = o.test(xt[:100], False, True)
    assert _rel(res0.get_mean(0), m0[:, 0]) < tol and np.abs(res0.get_mean_variance() - v0).max() < tol
    return g


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("kernel,scale,n,d,y_dim,frac", [("rbf", 0.5, 150, 1, 1, 1.0), ("rbf", 0.6, 200, 2, 2, 0.5), ("matern32", 0.8, 180, 2, 1, 0.7), ("matern32", 1.0, 120, 3, 2, 0.4),
                                                          ("rbf", 0.7, 90, 3, 1, 0.0)])
def test_noisy_input_gp_matches_oracle(gp, oracle, dtype, kernel, scale, n, d, y_dim, frac):
    _check(gp, oracle, dtype, kernel, scale, n, d, y_dim, frac, 700, seed=n + d)


def test_noisy_input_gp_many_test_points_and_larger_system(gp, oracle):
    """m = 400 + 2 * 400 = 1200 rows (ten 128-column panels of the blocked Cholesky), 9000 test points (two tiles)."""
    _check(gp, oracle, np.float64, "rbf", 0.3, 400, 2, 1, 1.0, 9000, seed=5)


def test_noisy_input_gp_reference_gtest(gp, oracle):
    """The reference's own test (test/gtest/test_noisy_input_gp.cpp:13-180): RBF l = 0.2, 100 samples of sin(2x) with gradients,
    noise 1e-4, 200 test points; its assertions mae < 1e-5, mae_grad < 1e-4 and the MAE values printed in its source."""
    n, t = 100, 200
    x, xt = np.linspace(0, 2 * np.pi, n)[:, None], np.linspace(0, 2 * np.pi, t)[:, None]
    y, gr = np.sin(2 * x[:, 0]), 2 * np.cos(2 * x[:, 0])
    g = gp.NoisyInputGaussianProcess(gp.NoisyInputGaussianProcess.Setting("rbf", 0.2, max_num_samples=n), np.float64)
    assert g.train(x, y, gr[:, None, None], 1e-4, 1e-4, 1e-4, 1)
    res = g.test(xt, True)
    mae = np.abs(res.get_mean(0) - np.sin(2 * xt[:, 0])).mean()
    mae_grad = np.abs(res.get_gradient(0)[0][:, 0] - 2 * np.cos(2 * xt[:, 0])).mean()
    assert mae < 1e-5 and mae_grad < 1e-4                                      # :179-180
    assert abs(mae - 4.1624286843223515e-06) / 4.16e-06 < 1e-5                 # :177
    assert abs(mae_grad - 7.139121709502966e-05) / 7.14e-05 < 1e-5
    # without gradient observations (:188-351): the gradient is still predicted
    g2 = gp.NoisyInputGaussianProcess(gp.NoisyInputGaussianProcess.Setting("rbf", 0.2, n, no_gradient_observation=True), np.float64)
    assert g2.train(x, y, None, 1e-4, 1e-4, None, 0)
    res2 = g2.test(xt, True)
    mae = np.abs(res2.get_mean(0) - np.sin(2 * xt[:, 0])).mean()
    mae_grad = np.abs(res2.get_gradient(0)[0][:, 0] - 2 * np.cos(2 * xt[:, 0])).mean()
    assert mae < 1e-4 and mae_grad < 0.0025                                    # :350-351
    assert abs(mae - 7.377464439757659e-05) / 7.38e-05 < 1e-5                  # :349
    assert abs(mae_grad - 0.0024347632450979033) / 2.43e-03 < 1e-5
    with pytest.raises(ValueError):
        g2.train(np.zeros((n + 1, 1)), np.zeros(n + 1), None, 1e-4, 1e-4, None, 0)  # max_num_samples, :711-714


def test_noisy_input_gp_limits(gp, oracle):
    rng = np.random.default_rng(3)
    x, y, grad, flag = _problem(rng, 40, 2, 1, 1.0)
    ou = gp.NoisyInputGaussianProcess(gp.NoisyInputGaussianProcess.Setting("ou", 0.5), np.float64)
    with pytest.raises(gp.ErlGpError):
        ou.train(x, y, grad, 1e-2, 1e-2, 1e-2, flag)  # OrnsteinUhlenbeck has no derivative at r = 0
    ou2 = gp.NoisyInputGaussianProcess(gp.NoisyInputGaussianProcess.Setting("ou", 0.5, no_gradient_observation=True), np.float64)
    assert ou2.train(x, y, None, 1e-2, 1e-2, None, 0)
    o = oracle.NoisyInputGp(oracle.OU, 0.5, True, np.float64)
    assert o.train(x, y, None, 1e-2, 1e-2, None, 0)
    xt = rng.uniform(-1, 1, (50, 2))
    m_r, _, v_r, _, _ = o.test(xt, False, True)
    res = ou2.test(xt, False)
    assert _rel(res.get_mean(0), m_r[:, 0]) < 1e-10 and np.abs(res.get_mean_variance() - v_r).max() < 1e-10
    with pytest.raises(gp.ErlGpError):
        ou2.test(xt, True).get_mean(0)
    bad = gp.NoisyInputGaussianProcess(gp.NoisyInputGaussianProcess.Setting("rbf", 0.5), np.float64)
    with pytest.raises(gp.ErlGpError):
        bad.train(rng.uniform(0, 1, (10, 4)), np.zeros(10), None, 1e-2, 1e-2, None, 0)  # x_dim > 3
