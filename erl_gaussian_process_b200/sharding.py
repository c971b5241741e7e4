"""Multi-GPU sharding of a stream of independent GPs.

The path has no exchange step: the per-partition loops of the reference
(src/lidar_gp_2d.cpp:366-392, src/range_sensor_gp_3d.cpp:334-360) have no cross-iteration dependency, so
GPs are split into contiguous ranges, one process per GPU computes its range through the C ABI, and the results
are gathered on the host (torch.distributed is used only for that gather - point-to-point receives straight into the slices of the
full arrays on rank 0; no NCCL collective is on the data path).

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


def _gather_rows(part: dict, row_ranges: dict, rank: int, world: int, group=None):
    """Host gather of per-rank arrays that are contiguous row ranges of full arrays every rank can size: rank 0 allocates the full
    arrays once and RECEIVES every other rank's rows straight into their slice (torch.distributed send / recv of zero-copy tensor
    views: no pickling, no concatenation); `row_ranges[key][r] = (begin, end)` of rank r along axis 0.  The point-to-point calls need
    a backend that moves host tensors (gloo); over an NCCL-only group the pickled `gather_object` is the fallback."""
    import torch
    import torch.distributed as dist

    if dist.get_backend(group) == "nccl":
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(part, gathered, dst=0, group=group)
        if rank != 0:
            return None
        return {k: np.concatenate([g[k] for g in gathered], axis=0) for k in part}
    if rank != 0:
        for k in sorted(part):
            a = np.ascontiguousarray(part[k])
            if a.size:
                dist.send(torch.from_numpy(a), dst=0, group=group)
        return None
    full = {}
    for k in sorted(part):
        total = row_ranges[k][world - 1][1]
        full[k] = np.empty((total,) + part[k].shape[1:], dtype=part[k].dtype)
        b0, b1 = row_ranges[k][0]
        full[k][b0:b1] = part[k]
    for r in range(1, world):
        for k in sorted(part):
            b0, b1 = row_ranges[k][r]
            if b1 > b0 and full[k][b0:b1].size:
                dist.recv(torch.from_numpy(full[k][b0:b1]), src=r, group=group)
    return full


def sharded_train_predict(compute, n_train, x, y, var, q_offsets, q_x, rank: int = 0, world: int = 1, group=None):
    """Run `compute(n_train, x, y, var, q_offsets, q_x) -> dict(mean, var, valid, info, alpha)` on this rank's shard and
    gather the shards on rank 0 (host gather).  Returns the full-batch dict on rank 0 and None elsewhere."""
    shard = shard_batch(rank, world, n_train, x, y, var, q_offsets, q_x)
    out = compute(shard["n_train"], shard["x"], shard["y"], shard["var"], shard["q_offsets"], shard["q_x"])
    part = {k: np.asarray(out[k]) for k in ("mean", "var", "valid", "info", "alpha")}
    if world == 1:
        return part
    # ranks hold contiguous, ascending GP ranges (and, in CSR order, contiguous query ranges): every rank can size every slice
    q_offsets = np.asarray(q_offsets, dtype=np.int64)
    gp_ranges = [shard_range(len(n_train), r, world) for r in range(world)]
    q_ranges = [(int(q_offsets[a]), int(q_offsets[b])) for a, b in gp_ranges]
    ranges = {"mean": q_ranges, "var": q_ranges, "valid": q_ranges, "info": gp_ranges, "alpha": gp_ranges}
    return _gather_rows(part, ranges, rank, world, group)


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
    t_ranges = [shard_range(x_test.shape[0], r, world) for r in range(world)]
    full = _gather_rows({"mean": part[0], "variance": part[1]}, {"mean": t_ranges, "variance": t_ranges}, rank, world, group)
    if full is None:
        return None
    return full["mean"], full["variance"]
