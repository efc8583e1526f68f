"""Multi-GPU plumbing: one process per GPU, torch.distributed for launch and the
single exchange step of the sharded matcher.

* detect + describe shards by frame: `shard_range` gives each rank a contiguous
  frame range; there is no collective on the data path.
* brute-force matching of a large train set shards the TRAIN rows: every rank
  computes its local top-k as packed keys (distance << 32 | global train index),
  the keys are all-gathered (NCCL on GPUs) and merged per query.  Unsigned key
  order is (distance, index), i.e. the reference's rule that the first minimum
  of a left-to-right scan wins (brute-force-matcher.cc:138-157).
"""
import numpy as np

KEY_NONE = np.uint64(0xFFFFFFFFFFFFFFFF)


def shard_range(n, rank, world):
    """Contiguous [begin, end) of `n` units for `rank` of `world` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def pack_keys(idx, dist, global_offset=0):
    """(idx, dist) int arrays [nq, k] -> uint64 keys; idx < 0 marks a missing neighbour."""
    idx = np.asarray(idx, np.int64)
    dist = np.asarray(dist, np.int64)
    keys = (dist.astype(np.uint64) << np.uint64(32)) | (idx + global_offset).astype(np.uint64)
    keys[idx < 0] = KEY_NONE
    return keys


def merge_keys_host(gathered, k):
    """Host statement of the merge rule: gathered uint64 [shards, nq, k] -> (idx, dist) [nq, k]."""
    g = np.asarray(gathered, np.uint64)
    s, nq, kk = g.shape
    allk = np.sort(g.transpose(1, 0, 2).reshape(nq, s * kk), axis=1)[:, :k]
    idx = (allk & np.uint64(0xFFFFFFFF)).astype(np.int64).astype(np.int32)
    dist = (allk >> np.uint64(32)).astype(np.int64).astype(np.int32)
    none = allk == KEY_NONE
    idx[none] = -1
    dist[none] = -1
    return idx, dist


def all_gather_keys(local_keys, group=None):
    """All-gather of per-rank key tensors [nq, k] (int64 view of the uint64 keys) -> [world, nq, k]."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = torch.empty((world,) + tuple(local_keys.shape), dtype=local_keys.dtype, device=local_keys.device)
    if local_keys.is_cuda:
        dist.all_gather_into_tensor(out, local_keys.contiguous(), group=group)
    else:  # gloo has no all_gather_into_tensor for all versions; use the list form
        parts = [torch.empty_like(local_keys) for _ in range(world)]
        dist.all_gather(parts, local_keys.contiguous(), group=group)
        out = torch.stack(parts)
    return out


def sharded_knn(matcher, query, train_shard, k, global_offset, group=None):
    """kNN of `query` against a train set sharded over the ranks of `group` (CUDA tensors).

    Every rank gets the full result: (idx [nq, k] int32 global train indices, dist [nq, k] int32)."""
    import torch
    import torch.distributed as dist
    nq = query.shape[0]
    keys = torch.empty((nq, k), dtype=torch.int64, device=query.device)
    matcher.knn_keys(query, train_shard, k, global_offset, keys)  # returns with the keys complete (it synchronises)
    gathered = all_gather_keys(keys, group)
    if gathered.is_cuda:
        # the collective is ordered on torch's stream only; the merge kernel runs on the context's stream
        torch.cuda.current_stream(gathered.device).synchronize()
    idx = torch.empty((nq, k), dtype=torch.int32, device=query.device)
    dst = torch.empty((nq, k), dtype=torch.int32, device=query.device)
    matcher.merge_keys(gathered, dist.get_world_size(group), nq, k, idx, dst)
    return idx, dst
