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
    # rank-major exchange used by the parity check of a gathered step: record j of rank r belongs to global proof r + j * world
    recs = b"".join(int(idx).to_bytes(4, "little") for idx in mine).ljust(4 * ((total + world - 1) // world), b"\xff")
    by_rank = bench.gather_by_rank(recs, rank, world, device="cpu")
    if rank == 0:
        for idx in range(total):
            assert int.from_bytes(by_rank[idx % world][4 * (idx // world):4 * (idx // world) + 4], "little") == idx
    else:
        assert by_rank is None
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


# ---- one MSM sharded by base range (BASELINE configs[4]); CPU stand-ins for the device calls ----------------------
def _msm_worker(rank, world, port, n, group, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import random
    import manta_rs_b200  # noqa: F401
    from manta_rs_b200 import sharded
    from helpers import cref, BLS12_381 as C
    rng = random.Random(5)
    ks = [rng.randrange(1, C.r) for _ in range(n)]
    sc = [rng.randrange(C.r) for _ in range(n)]
    bases = cref.fixed_base(group, ks)
    scalars = b"".join(s.to_bytes(32, "little") for s in sc)

    def local_msm(b, s, cnt):
        vals = [int.from_bytes(s[32 * i:32 * i + 32], "little") for i in range(cnt)]
        return cref.msm(group, b, vals, threads=1), 0.0

    def point_sum(points, cnt):
        return cref.msm(group, points, [1] * cnt, threads=1)

    out, _ = sharded.msm_sharded(group, bases, scalars, rank=rank, world=world, tensor_device="cpu",
                                 local_msm=local_msm, point_sum=point_sum)
    expect = cref.fixed_base(group, [sum(k * s for k, s in zip(ks, sc)) % C.r])
    assert out == expect, f"rank {rank}"
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_msm_world2():
    for group, n in ((1, 37), (2, 9), (1, 1)):
        mp.spawn(_msm_worker, args=(2, _free_port(), n, group, None), nprocs=2, join=True)


def test_shard_range_partitions():
    from manta_rs_b200 import sharded
    import manta_rs_b200  # noqa: F401
    for n in (0, 1, 5, 1 << 20):
        for world in (1, 2, 3, 8):
            edges = [sharded.shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
