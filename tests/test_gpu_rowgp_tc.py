"""GPU parity of the tcgen05 / TMEM fused train + predict kernel (erl_gp_rowgp_tc.cuh; FP32, n <= 128) against the oracle:
persistent CTAs walking many GPs (mbarrier phases carried across GPs), ragged training sets and query lists (0 queries,
more than the 128 that ride along with the factorisation), every covariance kernel and input dimension, an ill-conditioned
batch.  (The SASS evidence that the tensor path is tcgen05 - UTCHMMA / LDTM / STTM - is checked in tests/test_capi_exports.py,
where the object files are.)"""
import numpy as np
import pytest

from tests.util import err_mean, err_var, make_batch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gp():
    import erl_gaussian_process_b200 as m

    return m


@pytest.fixture(scope="module")
def tc_ctx(gp):
    """A context that routes the fused train + predict of n <= 128 through the tcgen05 kernel (the default is the mma.sync kernel)."""
    ctx = gp.Context(0)
    ctx.set_rowgp_tc(1)
    return ctx


def _run(gp, oracle, kernel, scale, batch, max_n, x_dim, min_num_samples=0, ctx=None):
    n_train, x, y, var, q_offsets, q_x = batch
    num_gps = len(n_train)
    launches0 = ctx.kernel_launches
    out = gp.BatchGp(num_gps, max_n, x_dim, kernel, scale, np.float32, ctx).train_predict(n_train, x, y, var, q_offsets, q_x, min_num_samples=min_num_samples)
    assert ctx.kernel_launches > launches0
    kid = oracle.KERNELS[kernel]
    nt_ref = np.where(n_train > min_num_samples, n_train, 0).astype(np.int32)
    ref32 = oracle.batched_train_predict(kid, scale, nt_ref, x, y, var, q_offsets, q_x)
    ref64 = oracle.batched_train_predict(kid, scale, nt_ref, x.astype(np.float64), y.astype(np.float64), var.astype(np.float64), q_offsets, q_x.astype(np.float64))
    trained = nt_ref > 0
    assert np.array_equal(out["info"] == 0, trained)
    qmask = np.repeat(trained, np.diff(q_offsets))
    assert np.array_equal(out["valid"], qmask)
    assert np.isnan(out["mean"][~qmask]).all() and np.isnan(out["var"][~qmask]).all()
    errs = {}
    for name, ref in (("f32", ref32), ("f64", ref64)):
        errs[name] = (err_mean(out["mean"][qmask], ref["mean"][qmask]), err_var(out["var"][qmask], ref["var"][qmask]))
        assert errs[name][0] < 1e-4 and errs[name][1] < 1e-4, (name, errs[name])  # north_star tolerance in float
    for g in np.flatnonzero(trained)[:: max(1, trained.sum() // 12)]:
        n = n_train[g]
        lg, lr = out["L"][g][:n, :n], ref64["L"][g][:n, :n]
        assert np.abs(np.triu(lg, 1)).max() == 0
        assert np.abs(lg - lr).max() / np.abs(lr).max() < 2e-5, g
        ag, ar = out["alpha"][g][:n], ref64["alpha"][g][:n]
        assert np.abs(ag - ar).max() / np.abs(ar).max() < 5e-3, g
    return out, errs


def test_tc_many_gps_per_cta(gp, oracle, tc_ctx):
    """2500 GPs on 296 persistent CTAs: ~8 GPs per CTA, n = 128, 128 queries each (the C4 shape)."""
    rng = np.random.default_rng(60)
    batch = make_batch(rng, 2500, 128, 3, np.float32, fixed_q=128)
    _run(gp, oracle, "matern32", 0.3, batch, 128, 3, ctx=tc_ctx)


@pytest.mark.parametrize("kernel,scale,x_dim", [("ou", 0.05, 1), ("matern32", 0.2, 2), ("rbf", 0.5, 3), ("matern32", 0.3, 3)])
def test_tc_ragged(gp, oracle, tc_ctx, kernel, scale, x_dim):
    """n from 0 to 128 (every panel count, padded last panels), 0 .. 400 queries per GP (tiles beyond the first 128 go through
    the mma.sync predict on the same shared-memory layout), min_num_samples gate."""
    rng = np.random.default_rng(61 + x_dim)
    batch = make_batch(rng, 900, 128, x_dim, np.float32, n_lo=0, n_hi=128, q_lo=0, q_hi=400)
    out, _ = _run(gp, oracle, kernel, scale, batch, 128, x_dim, min_num_samples=5, ctx=tc_ctx)
    assert (out["info"] == -1).any() and (out["info"] == 0).any()


@pytest.mark.parametrize("max_n", [16, 33, 64, 100])
def test_tc_small_capacities(gp, oracle, tc_ctx, max_n):
    rng = np.random.default_rng(70 + max_n)
    batch = make_batch(rng, 700, max_n, 2, np.float32, n_lo=1, n_hi=max_n, q_lo=0, q_hi=150)
    _run(gp, oracle, "matern32", 0.3, batch, max_n, 2, ctx=tc_ctx)


def test_tc_ill_conditioned(gp, oracle, tc_ctx):
    """cond(K) ~ 1e4 (pixel patch) at n = 128."""
    rng = np.random.default_rng(80)
    b, pr, pc = 400, 16, 8
    n = pr * pc
    r, c = np.meshgrid(np.arange(pr), np.arange(pc), indexing="ij")
    base = np.stack([r.ravel() * 1.6e-3, c.ravel() * 1.6e-3], axis=1)
    x = (base[None] + rng.uniform(-0.3, 0.3, (b, 1, 2))).astype(np.float32)
    y = (1.0 / np.sqrt(3.0 + rng.uniform(0, 2, (b, 1)) + 0.3 * np.sin(40 * x[..., 0]) * np.cos(55 * x[..., 1]))).astype(np.float32)
    var = np.full((b, n), 0.01, dtype=np.float32)
    n_train = np.full(b, n, dtype=np.int32)
    t = 128
    q_offsets = np.arange(b + 1, dtype=np.int64) * t
    lo, hi = x.min(axis=1, keepdims=True), x.max(axis=1, keepdims=True)
    q_x = (lo + (hi - lo) * rng.random((b, t, 2))).astype(np.float32).reshape(-1, 2)
    _run(gp, oracle, "matern32", 0.05, (n_train, x, y, var, q_offsets, q_x), n, 2, ctx=tc_ctx)


def test_tc_not_spd_and_deterministic(gp, tc_ctx):
    rng = np.random.default_rng(81)
    batch = list(make_batch(rng, 600, 128, 3, np.float32, n_lo=1, n_hi=128, q_lo=0, q_hi=200))
    batch[3][7, :] = -5.0  # K[i,i] = 1 + var < 0
    b = gp.BatchGp(600, 128, 3, "matern32", 0.3, np.float32, tc_ctx)
    ref = None
    for _ in range(3):
        out = b.train_predict(*batch)
        assert out["info"][7] > 0 and (np.delete(out["info"], 7) == 0).all()
        q0, q1 = batch[4][7], batch[4][8]
        assert not out["valid"][q0:q1].any() and out["valid"][:q0].all()
        cur = {k: np.array(out[k], copy=True) for k in ("mean", "var", "valid", "info", "alpha", "L")}
        if ref is None:
            ref = cur
        else:
            for k in ref:
                assert np.array_equal(ref[k], cur[k], equal_nan=True), k
