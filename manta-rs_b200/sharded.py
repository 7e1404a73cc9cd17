"""One large MSM sharded by base range across the GPUs of a node (BASELINE configs[4]: trusted-setup-sized G2 MSM).

Mirrors what `manta-trusted-setup` does with one `VariableBaseMSM::multi_scalar_mul` call over 2^19..2^20 bases
(`manta-trusted-setup/src/groth16/mpc.rs:367-381`, `kzg.rs:509-523`): rank r owns the contiguous
slice [lo_r, hi_r) of (bases, scalars), computes its partial sum with `mp_msm_g1/g2`, and the N affine partials (96 / 192
bytes each) are exchanged with ONE all_gather — the only data-path collective of the whole framework — and added with
`mp_points_sum_g1/g2`.  Group addition is associative, so the bytes equal those of the single-GPU MSM.
"""
from __future__ import annotations

import ctypes

POINT_BYTES = {1: 96, 2: 192}


def shard_range(n: int, rank: int, world: int):
    """Contiguous slice of rank `rank`: sizes differ by at most one, empty slices only when n < world."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _native_msm(group, device):
    from . import _native as nat
    lib = nat.lib()
    fn = lib.mp_msm_g1 if group == 1 else lib.mp_msm_g2

    def run(bases: bytes, scalars: bytes, n: int):
        out = ctypes.create_string_buffer(POINT_BYTES[group])
        ms = ctypes.c_float()
        nat.check(fn(device, bases, scalars, n, out, ctypes.byref(ms)))
        return out.raw, ms.value
    return run


def _native_sum(group, device):
    from . import _native as nat
    lib = nat.lib()
    fn = lib.mp_points_sum_g1 if group == 1 else lib.mp_points_sum_g2

    def run(points: bytes, n: int):
        out = ctypes.create_string_buffer(POINT_BYTES[group])
        nat.check(fn(device, points, n, out))
        return out.raw
    return run


def msm_sharded(group: int, bases: bytes, scalars: bytes, *, rank: int, world: int, device: int = 0, tensor_device="cuda",
                local_msm=None, point_sum=None):
    """Returns (result bytes on every rank, device ms of this rank's partial MSM).

    bases: n ark-uncompressed points; scalars: n x 32 bytes canonical little-endian.  `local_msm` / `point_sum` default to
    the CUDA library; the gloo test injects CPU stand-ins to exercise the sharding and the exchange without a GPU.
    """
    pb = POINT_BYTES[group]
    n = len(bases) // pb
    assert len(bases) == n * pb and len(scalars) == n * 32
    lo, hi = shard_range(n, rank, world)
    local_msm = local_msm or _native_msm(group, device)
    point_sum = point_sum or _native_sum(group, device)
    partial, ms = local_msm(bases[lo * pb:hi * pb], scalars[lo * 32:hi * 32], hi - lo)
    if world == 1:
        return partial, ms
    import torch
    import torch.distributed as dist
    mine = torch.frombuffer(bytearray(partial), dtype=torch.uint8).to(tensor_device)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    stacked = b"".join(bytes(p.cpu().numpy()) for p in parts)
    return point_sum(stacked, world), ms
