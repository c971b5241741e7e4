"""Multi-GPU sharding of a stream of independent GPs.

The path has no exchange step: the per-partition loops of the reference
(src/lidar_gp_2d.cpp:366-392, src/range_sensor_gp_3d.cpp:334-360) have no cross-iteration dependency, so
GPs are split into contiguous ranges, one process per GPU computes its range through the C ABI, and the results
are gathered on the host (torch.distributed is used only for that gather; no NCCL collective is on the data path).

`compute` is injected so that the host logic (range arithmetic, CSR re-basing of the per-GP query lists, gather
order) can be tested on CPU with world_size 2 over gloo (tests/test_sharding_gloo.py); on a GPU box it is
`BatchGp.train_predict`.
"""
from __future__ import annotations

import numpy as np


def shard_range(num_gps: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous GP range [begin, end) of `rank`: sizes differ by at most one, lower ranks take the remainder."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} / world {world}")
    base, rem = divmod(int(num_gps), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_batch(rank: int, world: int, n_train, x, y, var, q_offsets, q_x):
    """Slice one rank's GPs and their queries out of a batch (CSR query offsets re-based to 0)."""
    b = len(n_train)
    g0, g1 = shard_range(b, rank, world)
    q_offsets = np.asarray(q_offsets, dtype=np.int64)
    t0, t1 = int(q_offsets[g0]), int(q_offsets[g1])
    return dict(gp_begin=g0, gp_end=g1, q_begin=t0, q_end=t1, n_train=np.ascontiguousarray(n_train[g0:g1]), x=np.ascontiguousarray(x[g0:g1]), y=np.ascontiguousarray(y[g0:g1]),
                var=np.ascontiguousarray(var[g0:g1]), q_offsets=(q_offsets[g0:g1 + 1] - t0).astype(np.int64), q_x=np.ascontiguousarray(q_x[t0:t1]))


def sharded_train_predict(compute, n_train, x, y, var, q_offsets, q_x, rank: int = 0, world: int = 1, group=None):
    """Run `compute(n_train, x, y, var, q_offsets, q_x) -> dict(mean, var, valid, info, alpha)` on this rank's shard and
    gather the shards on rank 0 (host gather).  Returns the full-batch dict on rank 0 and None elsewhere."""
    shard = shard_batch(rank, world, n_train, x, y, var, q_offsets, q_x)
    out = compute(shard["n_train"], shard["x"], shard["y"], shard["var"], shard["q_offsets"], shard["q_x"])
    part = {k: np.asarray(out[k]) for k in ("mean", "var", "valid", "info", "alpha")}
    if world == 1:
        return part
    import torch.distributed as dist

    gathered = [None] * world if rank == 0 else None
    dist.gather_object(part, gathered, dst=0, group=group)
    if rank != 0:
        return None
    # ranks hold contiguous, ascending GP ranges: concatenation in rank order restores the batch order
    return {k: np.concatenate([g[k] for g in gathered], axis=0) for k in part}


def sharded_dense_predict(predict, x_test, rank: int = 0, world: int = 1, group=None):
    """Dense VanillaGaussianProcess predict over several GPUs: every rank holds a replica of the
    trained GP (L, alpha, x_train: the factorisation itself is "replicas only"), the test points are split into contiguous
    ranges, `predict(x_test_slice) -> (mean, variance)` runs on this rank's slice and the slices are gathered on rank 0 in
    rank order (host gather, no data-path collective).  Returns (mean, variance) on rank 0 and None elsewhere."""
    x_test = np.asarray(x_test)
    t0, t1 = shard_range(x_test.shape[0], rank, world)
    mean, variance = predict(np.ascontiguousarray(x_test[t0:t1]))
    part = (np.asarray(mean), np.asarray(variance))
    if world == 1:
        return part
    import torch.distributed as dist

    gathered = [None] * world if rank == 0 else None
    dist.gather_object(part, gathered, dst=0, group=group)
    if rank != 0:
        return None
    return np.concatenate([g[0] for g in gathered], axis=0), np.concatenate([g[1] for g in gathered], axis=0)
