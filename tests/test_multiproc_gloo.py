"""World-size-2 `gloo` test of the multi-rank plumbing of bench.py: round-robin sharding of a batch of proofs
 and the gather of proof bytes to rank 0 — no GPU, no data-path collective."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT

sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, total):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    mine = bench.shard_indices(total, rank, world)
    # fake 192-byte "proofs" that encode their global index
    local = torch.zeros((len(mine), 192), dtype=torch.uint8)
    for j, idx in enumerate(mine):
        local[j, 0] = idx % 256
        local[j, 1] = idx // 256
        local[j, 191] = rank
    gathered = bench.gather_proofs(local, total, rank, world, device="cpu")
    t = bench.max_over_ranks(float(rank + 1), device="cpu")
    assert t == float(world)
    if rank == 0:
        assert gathered.shape == (total, 192)
        for idx in range(total):
            assert int(gathered[idx, 0]) + 256 * int(gathered[idx, 1]) == idx
            assert int(gathered[idx, 191]) == idx % world
    else:
        assert gathered is None
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    port = _free_port()
    mp.spawn(_worker, args=(2, port, 11), nprocs=2, join=True)


def test_shard_indices_cover_batch():
    import bench
    for total in (0, 1, 7, 128):
        for world in (1, 2, 4, 8):
            seen = sorted(i for r in range(world) for i in bench.shard_indices(total, r, world))
            assert seen == list(range(total))
