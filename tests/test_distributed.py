"""N>1 host logic on CPU: world_size-2 gloo runs of the frame sharding and of the sharded
matcher's exchange + merge (local top-k keys come from the oracle here; on GPUs they come from
brisk_hamming_knn_keys and the merge from brisk_knn_merge_keys)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ethzasl_brisk_b200 import distributed as bd
from ethzasl_brisk_b200.synthetic import random_descriptors


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, k, result_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import restate
    q = random_descriptors(64, 64, 5)
    t = random_descriptors(1001, 64, 6)
    t[10] = q[7]
    t[900] = q[7]  # tie across shards: the lower global index must win
    b, e = bd.shard_range(len(t), rank, world)
    idx, dst = restate.knn(q, t[b:e], k)
    keys = torch.from_numpy(bd.pack_keys(idx, dst, b).view(np.int64))
    gathered = bd.all_gather_keys(keys).numpy().view(np.uint64)
    midx, mdist = bd.merge_keys_host(gathered, k)
    ridx, rdist = restate.knn(q, t, k)
    ok = np.array_equal(midx, ridx) and np.array_equal(mdist, rdist) and midx[7, 0] == 10 and midx[7, 1] == 900
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        open(result_path, "w").write(str(int(flag.item())))
    dist.destroy_process_group()


def test_sharded_knn_exchange_and_merge_gloo(tmp_path):
    result = tmp_path / "ok.txt"
    mp.spawn(_worker, args=(2, _free_port(), 2, str(result)), nprocs=2, join=True)
    assert result.read_text() == "1"


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 1024, 1025):
        for world in (1, 2, 3, 8):
            ranges = [bd.shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in ranges]
            assert max(sizes) - min(sizes) <= 1


def test_merge_keys_handles_missing_neighbours():
    keys = np.full((2, 3, 2), bd.KEY_NONE, np.uint64)
    keys[0, 0] = bd.pack_keys(np.array([[4, -1]]), np.array([[9, -1]]))[0]
    keys[1, 0] = bd.pack_keys(np.array([[2, 5]]), np.array([[9, 11]]), 100)[0]
    idx, dst = bd.merge_keys_host(keys, 2)
    assert idx[0].tolist() == [4, 102] and dst[0].tolist() == [9, 9]
    assert idx[1].tolist() == [-1, -1] and dst[2].tolist() == [-1, -1]
