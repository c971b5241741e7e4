"""GPU parity: one-CTA-per-GP batched train / predict and the fused Gram kernels vs the oracle.
Everything goes through the C ABI (ctypes)."""
import numpy as np
import pytest

from oracle import oracle_np as onp
from tests.util import TOL, err_mean, err_var, make_batch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gp():
    import erl_gaussian_process_b200 as m

    return m


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("kernel", ["ou", "matern32", "rbf"])
@pytest.mark.parametrize("x_dim", [1, 2, 3])
def test_gram_matches_oracle(gp, oracle, dtype, kernel, x_dim):
    rng = np.random.default_rng(11)
    n, t = 300, 517  # ragged vs the 128 x 32 tile
    x = rng.uniform(-1, 1, (n, x_dim)).astype(dtype)
    xt = rng.uniform(-1, 1, (t, x_dim)).astype(dtype)
    var = rng.uniform(0.001, 0.1, n).astype(dtype)
    scale = 0.3
    tol = 2e-6 if dtype == np.float32 else 1e-13
    k = gp.compute_ktrain(kernel, scale, x, var)
    k_ref = oracle.gram_train(oracle.KERNELS[kernel], scale, x, var)
    assert np.abs(k - k_ref).max() < tol
    assert np.array_equal(np.diag(k), (dtype(1) + var).astype(dtype))  # K[i,i] = 1 + var[i] exactly
    kt = gp.compute_ktest(kernel, scale, x, xt)
    kt_ref = oracle.gram_test(oracle.KERNELS[kernel], scale, x, xt)
    assert kt.shape == (n, t)
    assert np.abs(kt - kt_ref).max() < tol


def _check_batch(gp, oracle, dtype, kernel, scale, num_gps, max_n, x_dim, seed, min_num_samples=0, **kw):
    rng = np.random.default_rng(seed)
    n_train, x, y, var, q_offsets, q_x = make_batch(rng, num_gps, max_n, x_dim, dtype, **kw)
    b = gp.BatchGp(num_gps, max_n, x_dim, kernel, scale, dtype)
    out = b.train_predict(n_train, x, y, var, q_offsets, q_x, min_num_samples=min_num_samples)
    kid = oracle.KERNELS[kernel]
    nt_ref = np.where(n_train > min_num_samples, n_train, 0).astype(np.int32)  # the reference's `cnt > min` gate
    ref = oracle.batched_train_predict(kid, scale, nt_ref, x, y, var, q_offsets, q_x)
    ref64 = oracle.batched_train_predict(kid, scale, nt_ref, x.astype(np.float64), y.astype(np.float64), var.astype(np.float64), q_offsets, q_x.astype(np.float64))
    tol = TOL[np.dtype(dtype)]
    trained = nt_ref > 0
    assert np.array_equal(out["info"] == 0, trained), "trained flags differ"
    assert np.all(out["info"][~trained] == -1)
    qmask = np.repeat(trained, np.diff(q_offsets))
    assert np.array_equal(out["valid"], qmask)
    assert np.isnan(out["mean"][~qmask]).all() and np.isnan(out["var"][~qmask]).all(), "outputs of untrained GPs must stay untouched"
    if qmask.any():
        for r, name in ((ref, "oracle"), (ref64, "oracle-f64")):
            em = err_mean(out["mean"][qmask], r["mean"][qmask])
            ev = err_var(out["var"][qmask], r["var"][qmask])
            assert em < tol, f"mean vs {name}: {em:.3e}"
            assert ev < tol, f"var vs {name}: {ev:.3e}"
    # L / alpha materialisation
    ltol = 2e-5 if dtype == np.float32 else 1e-11
    for g in np.flatnonzero(trained)[:8]:
        n = n_train[g]
        lg, lr = out["L"][g][:n, :n], ref64["L"][g][:n, :n]
        assert np.abs(np.triu(lg, 1)).max() == 0, "strict upper triangle of L must be zero"
        assert np.abs(lg - lr).max() / np.abs(lr).max() < ltol
        # alpha = K^-1 y carries cond(K): the bar is the measured distance between two correct implementations at
        # this precision (float: the FP32 port vs the FP64 port; double: the port vs LAPACK potrf / potrs), times 10
        ag, ar = out["alpha"][g][:n], ref64["alpha"][g][:n]
        if dtype == np.float32:
            other = ref["alpha"][g][:n]
        else:
            xg, yg, vg = (np.asarray(v[g][:n], dtype=np.float64) for v in (x, y, var))
            other = onp.vanilla_train(kid, scale, xg, yg, vg)[1]
        floor = np.abs(other - ar).max() / np.abs(ar).max()
        # (float: the 3xTF32 products of the row-GP kernel carry a unit round-off of ~2^-20 against 2^-24 for the FP32 port's FMAs, so up
        # to 16x the port's own distance from the FP64 answer is expected - measured 13x on the 1-D OU case; the bar is 32x)
        assert np.abs(ag - ar).max() / np.abs(ar).max() < max((32 if dtype == np.float32 else 10) * floor, 1e-5 if dtype == np.float32 else 1e-13)
    return out


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_batch_c4_shape(gp, oracle, dtype):
    # BASELINE config 4 at reduced batch: n = 128, 3-D inputs, Matern32 l = 0.3, 128 test points per GP
    _check_batch(gp, oracle, dtype, "matern32", 0.3, 96, 128, 3, seed=6, fixed_q=128)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("max_n,x_dim,kernel,scale", [(64, 1, "ou", 0.05), (43, 1, "ou", 0.05), (100, 2, "matern32", 0.2), (192, 2, "matern32", 0.3), (17, 3, "rbf", 0.5)])
def test_batch_ragged(gp, oracle, dtype, max_n, x_dim, kernel, scale):
    # ragged n (including 0 and tiny), ragged query counts (including 0 and > one tile)
    _check_batch(gp, oracle, dtype, kernel, scale, 37, max_n, x_dim, seed=max_n, n_lo=0, n_hi=max_n, q_lo=0, q_hi=300)


def test_batch_max_n_256_f32(gp, oracle):
    _check_batch(gp, oracle, np.float32, "matern32", 0.3, 12, 256, 2, seed=3, n_lo=200, n_hi=256, q_lo=1, q_hi=200)
    _check_batch(gp, oracle, np.float32, "ou", 0.1, 9, 250, 1, seed=4, n_lo=130, n_hi=250, q_lo=0, q_hi=150)
    _check_batch(gp, oracle, np.float32, "matern32", 0.4, 7, 256, 3, seed=5, n_lo=256, n_hi=256, q_lo=64, q_hi=64)


def test_batch_min_num_samples_gate(gp, oracle):
    # RangeSensorGaussianProcess3D trains iff cnt > min_num_samples_per_group (src/range_sensor_gp_3d.cpp:358)
    out = _check_batch(gp, oracle, np.float32, "matern32", 0.3, 50, 96, 2, seed=9, min_num_samples=32, n_lo=20, n_hi=45, q_lo=1, q_hi=20)
    assert (out["info"] == -1).any() and (out["info"] == 0).any()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("max_n,x_dim,kernel,scale,num_gps", [(300, 2, "matern32", 0.3, 10), (257, 1, "ou", 0.05, 7), (484, 2, "matern32", 0.2, 6), (700, 3, "matern32", 0.4, 3)])
def test_batch_large_n(gp, oracle, dtype, max_n, x_dim, kernel, scale, num_gps):
    """Partition GPs beyond the shared-memory kernels (n > 256 float / 192 double; src/range_sensor_gp_3d.cpp:213 puts no cap on
    row_group_size * col_group_size): L resident in HBM / L2, blocked Cholesky per CTA (erl_gp_largegp.cu).  Ragged sizes incl. 0."""
    _check_batch(gp, oracle, dtype, kernel, scale, num_gps, max_n, x_dim, seed=max_n, n_lo=0, n_hi=max_n, q_lo=0, q_hi=200)
    _check_batch(gp, oracle, dtype, kernel, scale, 3, max_n, x_dim, seed=max_n + 1, n_lo=max_n, n_hi=max_n, q_lo=65, q_hi=130)


def test_batch_large_n_gate_and_not_spd(gp, oracle):
    out = _check_batch(gp, oracle, np.float64, "matern32", 0.3, 12, 320, 2, seed=77, min_num_samples=200, n_lo=150, n_hi=320, q_lo=1, q_hi=40)
    assert (out["info"] == -1).any() and (out["info"] == 0).any()
    rng = np.random.default_rng(0)
    n_train, x, y, var, q_offsets, q_x = make_batch(rng, 3, 300, 2, np.float32, fixed_q=8)
    var[1, 40] = -5.0  # K[40][40] = 1 + var < 0: LLT fails at column 41
    res = gp.BatchGp(3, 300, 2, "rbf", 0.5, np.float32).train_predict(n_train, x, y, var, q_offsets, q_x)
    assert res["info"][1] > 0 and (res["info"][[0, 2]] == 0).all()
    assert not res["valid"][8:16].any() and res["valid"][:8].all() and res["valid"][16:].all()


def test_batch_limits(gp):
    with pytest.raises(gp.ErlGpError):
        gp.BatchGp(4, 2049, 2, "ou", 1.0, np.float32)
    with pytest.raises(gp.ErlGpError):
        gp.BatchGp(4, 2049, 2, "ou", 1.0, np.float64)
    with pytest.raises(gp.ErlGpError):
        gp.BatchGp(4, 64, 4, "ou", 1.0, np.float32)


def test_batch_not_spd_reports_info(gp):
    rng = np.random.default_rng(0)
    n_train, x, y, var, q_offsets, q_x = make_batch(rng, 4, 32, 2, np.float32, fixed_q=8)
    var[1, :] = -5.0  # K[i,i] = 1 + var < 0: not positive definite
    out = gp.BatchGp(4, 32, 2, "rbf", 0.5, np.float32).train_predict(n_train, x, y, var, q_offsets, q_x)
    assert out["info"][1] > 0 and (out["info"][[0, 2, 3]] == 0).all()
    assert not out["valid"][8:16].any() and out["valid"][:8].all()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("max_n,x_dim,kernel,scale", [(64, 1, "ou", 0.05), (192, 2, "matern32", 0.3)])
def test_batch_device_path_train_then_predict(gp, oracle, dtype, max_n, x_dim, kernel, scale):
    """Device-resident buffers: train kernel, then the predict-only kernel (L reloaded from HBM) with scatter index."""
    import torch

    rng = np.random.default_rng(21)
    num_gps = 24
    n_train, x, y, var, q_offsets, q_x = make_batch(rng, num_gps, max_n, x_dim, dtype, n_lo=max_n // 2, n_hi=max_n, q_lo=100, q_hi=900)
    t = q_x.shape[0]
    b = gp.BatchGp(num_gps, max_n, x_dim, kernel, scale, dtype)
    b.upload(n_train, x, y, var)
    b.train_dev(write_l=True)
    perm = rng.permutation(t).astype(np.int32)
    dev = torch.device("cuda:0")
    d_off = torch.from_numpy(q_offsets).to(dev)
    d_qx = torch.from_numpy(q_x).to(dev)
    d_perm = torch.from_numpy(perm).to(dev)
    d_mean = torch.full((t,), float("nan"), dtype=d_qx.dtype, device=dev)
    d_var = torch.full((t,), float("nan"), dtype=d_qx.dtype, device=dev)
    d_valid = torch.zeros(t, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    b.predict_dev(d_off, d_qx, t, d_mean, d_var, d_valid, q_out_index=d_perm)
    b.ctx.synchronize()
    ref = oracle.batched_train_predict({"ou": oracle.OU, "matern32": oracle.MATERN32}[kernel], scale, n_train, x, y, var, q_offsets, q_x)
    mean = d_mean.cpu().numpy()
    varo = d_var.cpu().numpy()
    assert d_valid.cpu().numpy().all()
    tol = TOL[np.dtype(dtype)]
    assert err_mean(mean[perm], ref["mean"]) < tol
    assert err_var(varo[perm], ref["var"]) < tol


@pytest.mark.parametrize("max_n", [128, 192])
def test_batch_is_deterministic(gp, max_n):
    """Warp-level hand-offs inside the row-GP kernel (pivot tile through shared memory, Dinv from lanes 16-31): every output
    must be bit-identical from run to run."""
    rng = np.random.default_rng(17)
    batch = make_batch(rng, 300, max_n, 3, np.float32, n_lo=1, n_hi=max_n, q_lo=0, q_hi=150)
    b = gp.BatchGp(300, max_n, 3, "matern32", 0.3, np.float32)
    ref = None
    for _ in range(3):
        out = b.train_predict(*batch)
        cur = {k: np.array(out[k], copy=True) for k in ("mean", "var", "valid", "info", "alpha")}
        if ref is None:
            ref = cur
        else:
            for k in ref:
                assert np.array_equal(ref[k], cur[k], equal_nan=True), k


@pytest.mark.parametrize("dtype,max_n", [(np.float32, 128), (np.float64, 64)])
def test_multi_device_single_process(gp, oracle, dtype, max_n):
    """erl_gp_batch_train_predict_multi_*: one process, one pipeline (host thread + streams) per device, contiguous GP ranges,
    every device writes straight into the caller's arrays.  Runs with two pipelines on device 0 everywhere and over every
    visible GPU where there is more than one; the result must be bit-identical to the single-batch call (the ranges are
    independent) and within tolerance of the oracle; ragged training sets / query lists, untrained GPs."""
    rng = np.random.default_rng(91)
    num_gps, x_dim = 2311, 3
    batch = make_batch(rng, num_gps, max_n, x_dim, dtype, n_lo=0, n_hi=max_n, q_lo=0, q_hi=200)
    n_train, x, y, var, q_offsets, q_x = batch
    single = gp.BatchGp(num_gps, max_n, x_dim, "matern32", 0.3, dtype).train_predict(*batch, min_num_samples=3, want_l=True)
    device_sets = [[0, 0], [0, 0, 0]]
    if gp._capi.device_count() > 1:
        device_sets.append(list(range(gp._capi.device_count())))
    for devices in device_sets:
        multi = gp.MultiDeviceBatchGp(num_gps, max_n, x_dim, "matern32", 0.3, dtype, devices=devices)
        assert sum(multi.counts) == num_gps and max(multi.counts) - min(multi.counts) <= 1
        out = multi.train_predict(*batch, min_num_samples=3, want_l=True)
        for k in ("info", "valid", "mean", "var", "alpha", "L"):
            assert np.array_equal(single[k], out[k], equal_nan=True), (devices, k)
    kid = oracle.KERNELS["matern32"]
    nt_ref = np.where(n_train > 3, n_train, 0).astype(np.int32)
    ref = oracle.batched_train_predict(kid, 0.3, nt_ref, x, y, var, q_offsets, q_x)
    qmask = np.repeat(nt_ref > 0, np.diff(q_offsets))
    tol = TOL[np.dtype(dtype)]
    assert err_mean(out["mean"][qmask], ref["mean"][qmask]) < tol and err_var(out["var"][qmask], ref["var"][qmask]) < tol
    with pytest.raises(gp.ErlGpError):  # fewer GPs than devices
        gp.MultiDeviceBatchGp(1, max_n, x_dim, "matern32", 0.3, dtype, devices=[0, 0])
