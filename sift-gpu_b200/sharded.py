"""Cross-device batching for the driver's matcher (reference src/main.cpp:25-40).

The reference matches one query set against one scene set on the CPU.  Two ways to spread that over the GPUs of a box,
one process per GPU:

  * query-sharded: every rank holds the whole train (scene) set and a contiguous slice of the query rows.  Rows of the
    result are independent, so there is no exchange at all (`shard_rows` + `Sift.match_knn2`).
  * train-sharded: the train set is too large for one device, or arrives already spread over the devices that described
    it.  Every rank matches ALL queries against its own train slice -- the exact or the tensor-core kernel, through the
    C ABI -- and the ranks then exchange their two best candidates per query: 16 bytes per query and rank, one
    all_gather.  `merge_knn2` keeps the two globally best with the matcher's own order (ascending distance, exact ties ->
    lowest train index), and the ratio test runs on the merged pair.  The global best two are always among the local best
    two of some rank, so indices and distances equal a single-device match of the whole train set.

The exchange is the only collective on this path and it is a real one (a top-2 reduction across train shards); the
detect+describe path itself stays collective-free.
"""
from __future__ import annotations

import numpy as np


def shard_rows(n_rows: int, world: int, rank: int):
    """Contiguous block [lo, hi) of n_rows for `rank` of `world` (sizes differ by at most one row)."""
    base, rem = divmod(n_rows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def merge_knn2(dist_parts, idx_parts):
    """Merge per-shard knn-2 results.

    dist_parts / idx_parts: sequences of [nq, 2] float32 / int32 arrays, one per train shard, indices already GLOBAL
    (local index + shard offset), missing neighbours marked idx -1 / dist +inf (a shard with fewer than two rows).
    Returns (idx [nq, 2] int32, dist [nq, 2] float32): ascending distance, exact ties -> lowest train index."""
    d = np.concatenate([np.asarray(p, dtype=np.float32) for p in dist_parts], axis=1)
    i = np.concatenate([np.asarray(p, dtype=np.int32) for p in idx_parts], axis=1)
    missing = i < 0
    d = np.where(missing, np.float32(np.inf), d)
    tie = np.where(missing, np.iinfo(np.int32).max, i).astype(np.int64)
    # lexicographic (distance, index): stable sort by index first, then by distance
    order = np.argsort(tie, axis=1, kind="stable")
    d1 = np.take_along_axis(d, order, axis=1)
    order2 = np.argsort(d1, axis=1, kind="stable")
    pick = np.take_along_axis(order, order2, axis=1)[:, :2]
    return np.take_along_axis(i, pick, axis=1).astype(np.int32), np.take_along_axis(d, pick, axis=1).astype(np.float32)


def ratio_test(dist, idx, ratio: float = 0.86):
    """m1.distance <= ratio * m2.distance (src/main.cpp:38), evaluated in double like the C ABI; rows with fewer than two
    matches are skipped (src/main.cpp:32)."""
    dist = np.asarray(dist, dtype=np.float64)
    return (np.asarray(idx)[:, 1] >= 0) & (dist[:, 0] <= ratio * dist[:, 1])


def match_knn2_train_sharded(local_match, query, train_local, train_offset: int, *, ratio: float = 0.86, all_gather=None):
    """knn-2 + ratio test of `query` against a train set spread over the ranks.

    local_match(query, train_local) -> (idx [nq, 2], dist [nq, 2], ...): this rank's matcher on its own slice (e.g.
    `lambda q, t: sift.match_knn2(q, t, norm, ratio, tensor_cores=True)`); train_offset: global index of the slice's first
    row; all_gather(array) -> list of every rank's array in rank order (None: single rank).
    Returns (idx, dist, good) identical on every rank."""
    query = np.ascontiguousarray(query, dtype=np.float32)
    if len(train_local):
        res = local_match(query, np.ascontiguousarray(train_local, dtype=np.float32))
        idx, dist = np.array(res[0], dtype=np.int32), np.array(res[1], dtype=np.float32)
    else:
        idx = np.full((len(query), 2), -1, np.int32)
        dist = np.full((len(query), 2), np.inf, np.float32)
    idx = np.where(idx >= 0, idx + np.int32(train_offset), np.int32(-1)).astype(np.int32)
    packed = np.concatenate([dist.view(np.int32), idx], axis=1)  # one [nq, 4] int32 block per rank: a single exchange
    parts = all_gather(packed) if all_gather is not None else [packed]
    dists = [np.ascontiguousarray(p[:, :2]).view(np.float32) for p in parts]
    idxs = [np.ascontiguousarray(p[:, 2:]) for p in parts]
    midx, mdist = merge_knn2(dists, idxs)
    return midx, mdist, ratio_test(mdist, midx, ratio)


def torch_all_gather(group=None):
    """all_gather for `match_knn2_train_sharded` over torch.distributed (NCCL on the GPUs, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist

    def gather(arr: np.ndarray):
        world = dist.get_world_size(group)
        dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
        t = torch.from_numpy(np.ascontiguousarray(arr)).to(dev)
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t, group=group)
        return [o.cpu().numpy() for o in outs]

    return gather
