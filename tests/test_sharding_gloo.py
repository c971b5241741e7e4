"""N > 1 host logic on CPU: world_size 2 over gloo.  GP ranges, CSR re-basing and the host gather of
erl_gaussian_process_b200.sharding are exercised with the CPU oracle injected as the per-shard compute
(test-only use of the oracle: the product path calls the C ABI on each rank's GPU)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from erl_gaussian_process_b200 import sharding  # noqa: E402
from tests.util import make_batch  # noqa: E402


def test_shard_range_covers_everything():
    for b in (0, 1, 7, 24, 50_000):
        for world in (1, 2, 3, 8):
            ranges = [sharding.shard_range(b, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == b
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in ranges]
            assert max(sizes) - min(sizes) <= 1


def test_shard_batch_rebases_queries():
    rng = np.random.default_rng(3)
    batch = make_batch(rng, 11, 16, 2, np.float32, n_lo=3, q_lo=0, q_hi=9)
    seen_q = 0
    for r in range(3):
        s = sharding.shard_batch(r, 3, *batch)
        assert s["q_offsets"][0] == 0 and s["q_offsets"][-1] == len(s["q_x"])
        assert s["q_begin"] == seen_q
        seen_q = s["q_end"]
    assert seen_q == len(batch[5])


def _worker(rank, world, port, tmpdir):
    import torch.distributed as dist

    import oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(11)  # same batch on every rank
    batch = make_batch(rng, 13, 32, 3, np.float32, n_lo=5, q_lo=0, q_hi=12)

    def compute(n_train, x, y, var, q_offsets, q_x):
        out = oracle.batched_train_predict(oracle.MATERN32, 0.3, n_train, x, y, var, q_offsets, q_x)
        out.setdefault("valid", np.ones(len(q_x), dtype=bool))
        return out

    full = sharding.sharded_train_predict(compute, *batch, rank=rank, world=world)
    if rank == 0:
        ref = compute(*batch)
        for k in ("mean", "var", "info", "alpha"):
            np.testing.assert_array_equal(full[k], np.asarray(ref[k]), err_msg=k)
    else:
        assert full is None

    # dense VanillaGaussianProcess predict: replicas of the trained GP, test points split across ranks, host gather
    n, t = 60, 101
    xd = rng.uniform(-1, 1, (n, 2))
    yd = np.sin(3 * xd).sum(axis=1)
    vd = np.full(n, 1e-3)
    xt = rng.uniform(-1, 1, (t, 2))
    van = oracle.VanillaGp(oracle.MATERN32, 0.4, np.float64, max_num_samples=n)
    assert van.train(xd, yd, vd) == 0
    res = sharding.sharded_dense_predict(lambda xs: van.test(xs), xt, rank=rank, world=world)
    if rank == 0:
        m_ref, v_ref = van.test(xt)
        np.testing.assert_array_equal(res[0], m_ref)
        np.testing.assert_array_equal(res[1], v_ref)
        np.save(os.path.join(tmpdir, "ok.npy"), np.array([1]))
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp

    import oracle

    oracle.build()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok.npy")
