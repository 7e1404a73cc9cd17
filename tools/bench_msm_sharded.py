"""BASELINE configs[4] stress: ONE G2 (or G1) MSM over 2^20 bases, sharded by base range across the GPUs of a node.

    python tools/bench_msm_sharded.py [--group 2] [--log-n 20]                      # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_msm_sharded.py

Every rank builds the same (bases, scalars) from a fixed seed, computes the partial sum of its slice with mp_msm_g2, the
partials meet in one NCCL all_gather (192 B per rank) and are added with mp_points_sum_g2.  The result is checked
against the closed form (sum k_i s_i) * G computed by the CPU oracle's fixed-base routine.  Prints one JSON line.
"""
import argparse, ctypes, json, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--group", type=int, default=2, choices=[1, 2])
    ap.add_argument("--log-n", type=int, default=20)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    rank, world, local_rank = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    import torch
    import torch.distributed as dist
    import manta_rs_b200  # noqa: F401
    from manta_rs_b200 import _native as nat, sharded, workload as wl
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = nat.lib()
    pb = sharded.POINT_BYTES[args.group]
    n, uniq = 1 << args.log_n, 1 << 12
    rng = random.Random(2020)
    base_k = [rng.randrange(1, wl.FR_BLS12_381) for _ in range(uniq)]
    pts = ctypes.create_string_buffer(uniq * pb)
    gen = lib.mp_fixed_base_g1 if args.group == 1 else lib.mp_fixed_base_g2
    nat.check(gen(local_rank, nat.pack_scalars(base_k), uniq, pts))
    bases = pts.raw * (n // uniq)
    sc = bytearray(random.Random(args.log_n).randbytes(n * 32))
    sc[31::32] = bytes(b & 0x3F for b in sc[31::32])          # < 2^254 < r
    scalars = bytes(sc)
    best_total, best_part = 1e9, 1e9
    out = None
    for _ in range(args.reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out, part_ms = sharded.msm_sharded(args.group, bases, scalars, rank=rank, world=world, device=local_rank,
                                           tensor_device=torch.device("cuda", local_rank))
        torch.cuda.synchronize()
        total_ms = (time.perf_counter() - t0) * 1e3
        if world > 1:
            t = torch.tensor([part_ms, total_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            part_ms, total_ms = float(t[0]), float(t[1])
        best_total, best_part = min(best_total, total_ms), min(best_part, part_ms)
    if rank == 0:
        from oracle import cref   # checker only
        acc = [0] * uniq
        for i in range(n):
            acc[i % uniq] += int.from_bytes(scalars[32 * i:32 * i + 32], "little")
        tot = sum(k * a for k, a in zip(base_k, acc)) % wl.FR_BLS12_381
        ok = out == cref.fixed_base(args.group, [tot])
        print(json.dumps({"metric": f"G{args.group} MSM 2^{args.log_n} bases sharded by base range", "n_gpus": world,
                          "device_ms_partial_msm_max_over_ranks": best_part, "wall_ms_incl_h2d_and_gather": best_total,
                          "exchange_bytes_per_rank": pb, "parity": "closed form ok" if ok else "MISMATCH"}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
