#!/usr/bin/env python3
"""Benchmark of the B200 Groth16 proving path — BASELINE.json metric "PrivateTransfer Groth16 proofs/sec".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]
                    [--workload prove|single|msm_sweep|g2_stress] [--shape ...] [--dist U|R]

Workloads (BASELINE.json `configs`; the default is the one the metric is quoted on):
  prove      configs[1] shape at configs[3] batching (default): one step = one pass of `create_proof` (witness map + 5 MSMs +
             finish) over a batch of B synthetic proofs per GPU.  For N > 1 the driver launches one rank per GPU with torchrun;
             proofs are sharded round-robin, no data-path collective, the proof bytes are gathered to rank 0 over NCCL (weak
             scaling).  `--shape to_public` / `to_private` select the other circuits (configs[0], [4]), `--dist R` the witness
             distribution with 10 % booleans.  The line also carries the single-proof latency (configs[1]) of the same context.
  single     configs[1]: one proof per step through `mp_prove` with host buffers (latency).
  msm_sweep  configs[2]: stand-alone G1 MSM, 2^16 .. 2^--max-log points, bases resident, CPU oracle beside it up to 2^20.
  g2_stress  configs[4]: ONE 2^20-base G2 MSM sharded by base range over the ranks, one 192-byte all_gather.

  value    metric with inputs already resident in HBM (CUDA-event time of the kernels, max over ranks)
  e2e      the same through the C ABI with HOST buffers: pinned H2D of every input, kernels, D2H of the result (+ the
           gather at N > 1), wall clock between device synchronisations, max over ranks
  roofline dominant kernel against the measured integer-pipe rate
  cpu_baseline  the CPU oracle (C++ restatement of the reference's arkworks path) on this host, rank 0, N = 1 only
  parity   every proof of the last step against the known-trapdoor closed form and a sample against the full CPU oracle,
           on rank 0 AFTER the gather, at every N

`--impl reference` times the CPU restatement alone on all host cores (the reference's Rust toolchain is absent from this
image, so `oracle/_ref` cannot exist; DESIGN.md "reference arm").
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import ctypes
import json
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SHAPE = "private_transfer"     # the headline workload; --shape selects the other two circuits (BASELINE configs[0], [4])
DIST = "U"
KEY_SEED = 21
MADD_MULS = 11
FR = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
# wide MACs of one Fq product on the multiplier pipe: 288 IMAD.WIDE + 12 IMAD (m_i) = the 300 slots of the peak's denominator;
# the dedicated squaring needs 222 + 12 (fp.cuh) -> 0.78 of a product
SQR_AS_MUL = (222 + 12) / 300.0
# credited work per proof in Fq-multiplication equivalents = pairs x reference windows x 11 (x 3 in G2)
#                  name: (label, n, p, log_m, G1 pairs, G2 pairs, reference windows)
SHAPE_INFO = {
    "private_transfer": ("PrivateTransfer", 35175, 27, 16, 171031, 35174, 20),
    "to_public": ("ToPublic", 27945, 19, 15, 116581, 27944, 22),
    "to_private": ("ToPrivate", 8253, 13, 14, 41127, 8252, 24),
}
DTYPE = "u32 limbs (Fq 381-bit / Fr 255-bit Montgomery, integer pipe)"


def set_shape(name, dist="U"):
    global SHAPE, DIST, LABEL, WORKLOAD, CREDIT_G1_PER_PROOF, CREDIT_G2_PER_PROOF, CREDIT_PER_PROOF, NTT_BYTES_PER_PROOF
    label, n, p, log_m, g1_pairs, g2_pairs, windows = SHAPE_INFO[name]
    SHAPE, LABEL, DIST = name, label, dist
    WORKLOAD = f"{name} n={n} p={p} m=2^{log_m}" + (", witness distribution R (10 % booleans, 1 % < 2^128)" if dist == "R" else "")
    CREDIT_G1_PER_PROOF = g1_pairs * windows * MADD_MULS           # PrivateTransfer: 37.63 M
    CREDIT_G2_PER_PROOF = g2_pairs * windows * MADD_MULS * 3       # 23.21 M
    CREDIT_PER_PROOF = CREDIT_G1_PER_PROOF + CREDIT_G2_PER_PROOF   # 60.84 M
    NTT_BYTES_PER_PROOF = 7 * 2 * 32 * (1 << log_m)                # 28 MiB algorithmic


set_shape(SHAPE)


def ark_window(n: int) -> int:
    """ark-ec 0.3 `multi_scalar_mul` window rule: c = 3 if n < 32 else ceil(log2 n) * 69 / 100 + 2."""
    return 3 if n < 32 else ((n - 1).bit_length() * 69) // 100 + 2


def credited_msm_fq_muls(n: int, g2: bool = False) -> int:
    """N * ceil(255 / c_ref(N)) * 11 Fq-mul-equivalents (x 3 in G2)."""
    c = ark_window(n)
    return n * (-(-255 // c)) * MADD_MULS * (3 if g2 else 1)


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ---- multi-rank plumbing (exercised on CPU/gloo by tests/test_multiproc_gloo.py) ------------------------------
def shard_indices(total: int, rank: int, world: int):
    """Round-robin partition of proof indices: proof i -> rank i mod world."""
    return list(range(rank, total, world))


_GATHER_PERM = {}


def gather_proofs(local, total: int, rank: int, world: int, device="cuda", width: int = 192):
    """Gather [len(shard), width] uint8 rows (proof bytes) from every rank to rank 0 in global proof order."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local
    per = (total + world - 1) // world
    padded = torch.zeros((per, width), dtype=torch.uint8, device=device)
    padded[: local.shape[0]] = local.to(device)
    bufs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded)
    if rank != 0:
        return None
    key = (total, world, str(device))
    perm = _GATHER_PERM.get(key)
    if perm is None:
        # global proof i lives at row (i mod world) * per + i div world of the rank-major stack
        perm = torch.tensor([(i % world) * per + i // world for i in range(total)], device=device)
        _GATHER_PERM[key] = perm
    return torch.cat(bufs, dim=0)[perm]


def gather_by_rank(local_bytes: bytes, rank: int, world: int, device="cuda"):
    """All ranks contribute equally long byte strings; rank 0 gets the list indexed by rank (None elsewhere).  Kept apart
    from gather_proofs on purpose: the parity check maps (rank, local index) -> global proof itself."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return [local_bytes]
    mine = torch.frombuffer(bytearray(local_bytes), dtype=torch.uint8).to(device)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    if rank != 0:
        return None
    return [bytes(p.cpu().numpy()) for p in parts]


def max_over_ranks(x: float, device="cuda") -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ---- workload -----------------------------------------------------------------------------------------------------
def _assignment_bytes(seed):
    from manta_rs_b200 import workload as wl
    cs = _assignment_bytes.cs
    z = wl.make_assignment(cs, seed)
    return b"".join(int(v).to_bytes(32, "little") for v in z)


def make_assignments(cs, seeds, pool=True):
    """Packed canonical assignments (n x 32 bytes each), generated in forked workers before CUDA is touched."""
    _assignment_bytes.cs = cs
    workers = max(1, min(len(seeds), host_threads() // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1"))), 16))
    if not pool or workers == 1 or len(seeds) < 4:
        return [_assignment_bytes(s) for s in seeds]
    with cf.ProcessPoolExecutor(max_workers=workers, mp_context=mp.get_context("fork")) as ex:
        return list(ex.map(_assignment_bytes, seeds, chunksize=max(1, len(seeds) // (4 * workers))))


def randomness(seeds, modulus):
    """(r, s) per proof, drawn like `create_random_proof` from a ChaCha20Rng seeded with the proof index."""
    from manta_rs_b200.rng import ChaCha20Rng, field_rand
    rs, ss = [], []
    for seed in seeds:
        rng = ChaCha20Rng(int(seed).to_bytes(32, "little"))
        rs.append(field_rand(rng, modulus))
        ss.append(field_rand(rng, modulus))
    return rs, ss


def unpack_fr(buf: bytes):
    return [int.from_bytes(buf[i:i + 32], "little") for i in range(0, len(buf), 32)]


def random_scalars(n: int, seed: int) -> bytes:
    """n uniform Fr scalars, canonical little-endian 32 bytes each (ark `Fr::rand` shape: top bit shaved, reject >= r)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    out = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    out[:, 31] &= 0x7F
    mod_be = np.frombuffer(FR.to_bytes(32, "big"), dtype=np.uint8)
    while True:
        be = out[:, ::-1]
        diff = be.astype(np.int16) - mod_be.astype(np.int16)
        nz = diff != 0
        first = nz.argmax(axis=1)
        ge = ~nz.any(axis=1) | (diff[np.arange(n), first] > 0)
        bad = np.nonzero(ge)[0]
        if bad.size == 0:
            return out.tobytes()
        fresh = rng.integers(0, 256, size=(bad.size, 32), dtype=np.uint8)
        fresh[:, 31] &= 0x7F
        out[bad] = fresh


def dot_mod(a_bytes: bytes, b_bytes: bytes, modulus: int) -> int:
    """sum a_i b_i mod r over two packed arrays of 32-byte little-endian integers."""
    from operator import mul
    acc = 0
    step = 1 << 16
    for lo in range(0, len(a_bytes) // 32, step):
        a = unpack_fr(a_bytes[32 * lo:32 * (lo + step)])
        b = unpack_fr(b_bytes[32 * lo:32 * (lo + step)])
        acc = (acc + sum(map(mul, a, b))) % modulus
    return acc


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.idx = device_index
        self.rows = []
        self.proc = None
        self.thread = None
        self.t_begin = None

    def begin(self):
        """Marks the start of the timed region: nvidia-smi needs up to a second to come up on an 8-GPU box, so it is started
        before the warm-up steps and only the samples that arrive after this mark are reported."""
        self.t_begin = time.perf_counter()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return

        def pump():
            for line in self.proc.stdout:
                self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], 0.0, set()
        inside = [r for t, r in self.rows if self.t_begin is None or t >= self.t_begin]
        for r in inside or [r for _, r in self.rows[-2:]]:   # a region shorter than the sampling period: the last samples before it ends
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        load = [x for x in sm if x > 0.5 * mx] or sm
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def env_ranks():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0")))


def init_cuda(local_rank, world):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        print("bench.py: no CUDA device — the proving path has no CPU fallback", file=sys.stderr)
        sys.exit(2)
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    return torch, dist


def int_pipe_peak(nat, lib, device):
    wide, fqm = ctypes.c_double(), ctypes.c_double()
    nat.check(lib.mp_debug_int_pipe_rate(device, ctypes.byref(wide), ctypes.byref(fqm)))
    return wide.value / 300.0, fqm.value


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


# ---- reference arm ------------------------------------------------------------------------------------------------
REF_NOTE = ("C++ restatement of the reference's arkworks 0.3 CPU path (no Rust toolchain in this image, so the reference binary "
            "itself cannot be built); OpenMP over one flat (MSM, window) task list mirrors arkworks' optional `parallel` feature; the "
            "thread count is set explicitly (torchrun's OMP_NUM_THREADS=1 does not apply)")


def reference_line(args, metric, value, unit, dt, workload, threads, sample, higher=True, extra=None):
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps),
        "higher_is_better": higher, "scaling": "weak", "vs_baseline": None, "dtype": "u64 limbs (Fq 381-bit / Fr 255-bit Montgomery)",
        "data": "synthetic", "config": {"workload": workload},
        "cpu_baseline": {"value": value, "unit": unit, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": REF_NOTE,
    }
    if extra:
        line.update(extra)
    print(json.dumps(line), flush=True)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from manta_rs_b200 import workload as wl
    from oracle import cref
    threads = host_threads()
    if args.workload in ("msm_sweep", "g2_stress"):
        # bounded sample: one MSM of 2^16 (G1 sweep) / 2^14 (G2 stress) points per step on all host threads
        group = 1 if args.workload == "msm_sweep" else 2
        log_n = 16 if group == 1 else 14
        n = 1 << log_n
        bases = cref.fixed_base(group, unpack_fr(random_scalars(1 << 10, 7))) * (n >> 10)
        sc = random_scalars(n, 8)
        out = ctypes.create_string_buffer(96 * group)
        fn = cref.lib().oracle_msm_g1 if group == 1 else cref.lib().oracle_msm_g2
        for _ in range(args.warmup):
            fn(bases, sc, n, out, threads)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn(bases, sc, n, out, threads)
        dt = time.perf_counter() - t0
        credit = credited_msm_fq_muls(n, group == 2)
        value = credit * args.steps / dt / 1e9
        reference_line(args, f"G{group} MSM credited GFq-mul/s", value, "GFq-mul/s", dt,
                       f"stand-alone G{group} MSM, 2^{log_n} points per step on host cores (bounded sample of the {args.workload} workload)",
                       threads, f"one 2^{log_n}-point G{group} MSM per step, {threads} host threads")
        return 0
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import oracle_keygen
    cs = wl.make_shape(SHAPE, dist=DIST)
    pk, _ = oracle_keygen(cs, wl.sample_trapdoor(KEY_SEED))
    op = cref.OracleProver(pk, cs.p, cs.w, cs.a, cs.b, cs.c)
    zs = make_assignments(cs, list(range(args.warmup + args.steps)))
    rs, ss = randomness(list(range(len(zs))), cs.modulus)
    for i in range(args.warmup):
        op.prove(zs[i], rs[i], ss[i], threads=threads)
    t0 = time.perf_counter()
    for i in range(args.warmup, args.warmup + args.steps):
        op.prove(zs[i], rs[i], ss[i], threads=threads)
    dt = time.perf_counter() - t0
    value = args.steps / dt
    t0 = time.perf_counter()
    op.prove(zs[0], rs[0], ss[0], threads=1)
    one = 1.0 / (time.perf_counter() - t0)
    sample = f"1 {LABEL} proof per step (bounded sample of the batch), {threads} host threads (set explicitly)"
    reference_line(args, f"{LABEL} Groth16 proofs/sec", value, "proofs/s", dt, f"{WORKLOAD}, BLS12-381, 1 proof per step on host cores",
                   threads, sample, extra={"single_thread_value": one,
                                           "single_thread_note": "the reference as shipped enables no `parallel` feature: this is its configuration"})
    return 0


# ---- parity of a gathered step ---------------------------------------------------------------------------------------
def check_parity(cs, pk, rank, world, total, my_idx, z_list, rs, ss, gathered, sample, device):
    """Rank 0, after the gather: EVERY proof of the step against the known-trapdoor closed form (discrete logs computed by
    the rank that owns the assignment, exchanged rank-major, mapped to global order here), and `sample` proofs spread over
    all ranks against the full CPU oracle (rank 0 regenerates those assignments from their seeds).  Returns
    (parity string or None, cpu timings dict)."""
    from manta_rs_b200 import workload as wl
    from oracle import cref, trapdoor
    _, _, trap = trapdoor.key_scalars(cs, wl.sample_trapdoor(KEY_SEED))
    chk = trapdoor.TrapdoorChecker(cs, trap)
    mine = b"".join(b"".join(int(v).to_bytes(32, "little") for v in chk.scalars(unpack_fr(z), r, s)) for z, r, s in zip(z_list, rs, ss))
    by_rank = gather_by_rank(mine, rank, world, device)
    if rank != 0:
        return None, None
    triples = []
    for i in range(total):
        rec = by_rank[i % world][96 * (i // world):96 * (i // world + 1)]
        triples.append(tuple(unpack_fr(rec)))
    expect = trapdoor.TrapdoorChecker.bytes_from_scalars(triples)
    got = [bytes(gathered[192 * i:192 * (i + 1)]) for i in range(total)]
    bad = [i for i in range(total) if got[i] != expect[i]]
    # oracle sample
    timings = {}
    sample = min(sample, total)
    idx = sorted(set(int(round(k * (total - 1) / max(1, sample - 1))) for k in range(sample))) if sample else []
    obad = []
    if idx:
        threads = host_threads()
        op = cref.OracleProver(pk, cs.p, cs.w, cs.a, cs.b, cs.c)
        own = {g: j for j, g in enumerate(my_idx)}
        r2, s2 = randomness(idx, cs.modulus)
        t_all = []
        for k, g in enumerate(idx):
            z = z_list[own[g]] if g in own else make_assignments(cs, [g], pool=False)[0]
            t0 = time.perf_counter()
            ref = op.prove(z, r2[k], s2[k], threads=threads)
            t_all.append(time.perf_counter() - t0)
            if ref != got[g]:
                obad.append(g)
        timings = {"all_threads_s_per_proof": statistics.median(t_all), "threads": threads, "proofs": len(idx)}
        if world == 1:
            t0 = time.perf_counter()
            ref = op.prove(z_list[0], rs[0], ss[0], threads=1)
            timings["single_thread_s_per_proof"] = time.perf_counter() - t0
            if ref != got[my_idx[0]]:
                obad.append(my_idx[0])
        op.close()
    if bad or obad:
        return f"MISMATCH: trapdoor form {bad[:8]} ({len(bad)} of {total}), oracle {obad[:8]}", timings
    return (f"bit-exact: all {total} gathered proofs of the step vs the known-trapdoor closed form, {len(idx)} of them (spread over all "
            f"{world} rank(s)) vs the full CPU oracle"), timings


# ---- prove workload --------------------------------------------------------------------------------------------------
def run_prove(args):
    rank, world, local_rank = env_ranks()
    B = args.batch
    import manta_rs_b200  # noqa: F401
    from manta_rs_b200 import workload as wl
    cs = wl.make_shape(SHAPE, dist=DIST)
    total = B * world
    my_idx = shard_indices(total, rank, world)
    z_list = make_assignments(cs, my_idx)
    rs, ss = randomness(my_idx, cs.modulus)

    torch, dist = init_cuda(local_rank, world)
    from manta_rs_b200 import _native as nat, keygen, groth16 as g16
    lib = nat.lib()

    # ---- key + context (not timed)
    t_setup = time.perf_counter()
    g16.Groth16.device = local_rank
    pk = keygen.generate(cs, wl.sample_trapdoor(KEY_SEED), device=local_rank)
    ctx_obj = g16.ProvingContext.decode(pk)
    matrices = g16.R1CS.from_workload(cs, [1] + [0] * (cs.n - 1)).matrices
    ctx = ctx_obj.native(matrices, local_rank)
    # two batches in flight: the latency-bound tail and the host copies of one hide behind the kernels of the other
    NB = args.inflight
    batches = [ctypes.c_void_p() for _ in range(NB)]
    for i, bh in enumerate(batches):
        nat.check(lib.mp_batch_create_ex(ctx, B, 1 if i == 0 else 0, ctypes.byref(bh)))
        if args.no_g2_stream:
            nat.check(lib.mp_batch_set_overlap(bh, 0))
    batch = batches[0]
    n = cs.n
    z_host = torch.empty(B * n * 32, dtype=torch.uint8).pin_memory()
    z_host.numpy()[:] = memoryview(b"".join(z_list))
    r_host = torch.frombuffer(bytearray(nat.pack_scalars(rs)), dtype=torch.uint8).pin_memory()
    s_host = torch.frombuffer(bytearray(nat.pack_scalars(ss)), dtype=torch.uint8).pin_memory()
    out_hosts = [torch.empty(B * 192, dtype=torch.uint8).pin_memory() for _ in batches]
    setup_s = time.perf_counter() - t_setup

    def upload(bh):
        nat.check(lib.mp_batch_upload(bh, B, z_host.data_ptr(), r_host.data_ptr(), s_host.data_ptr()))

    def run(bh):
        ms = ctypes.c_float()
        nat.check(lib.mp_batch_run(bh, ctypes.byref(ms)))
        return ms.value

    def wait(bh):
        ms = ctypes.c_float()
        nat.check(lib.mp_batch_wait(bh, ctypes.byref(ms)))
        return ms.value

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput: W warm-up + K timed steps, assignments already in HBM
    for bh in batches:
        upload(bh)
    sampler = ClockSampler(local_rank)
    sampler.start()   # before the warm-up: the query process is up when the timed region starts
    for i in range(args.warmup):
        run(batches[i % NB])
    nphase = 8
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.begin()
    t0 = time.perf_counter()
    ev0.record()
    for k in range(args.steps):
        bh = batches[k % NB]
        if k >= NB:
            wait(bh)
        nat.check(lib.mp_batch_run_async(bh))
    for bh in batches:
        wait(bh)
    ev1.record()
    barrier()
    wall_s = time.perf_counter() - t0
    dev_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    # per-kernel (phase) times for the roofline: two extra steps with the two streams serialised, so that every
    # CUDA-event interval covers its kernels alone (not part of `value`)
    phase_acc = [0.0] * nphase
    nat.check(lib.mp_batch_set_overlap(batch, 0))
    serial_ms = 0.0
    for _ in range(2):
        serial_ms += run(batch)
        buf = (ctypes.c_float * nphase)()
        lib.mp_batch_phase_ms(batch, buf, nphase)
        for i in range(nphase):
            phase_acc[i] += buf[i] / 2
    nat.check(lib.mp_batch_set_overlap(batch, 0 if args.no_g2_stream else 1))
    launches = int(lib.mp_batch_kernel_launches(batch)) * args.steps
    dev_s = max_over_ranks(dev_ms * 1e-3)
    wall_s = max_over_ranks(wall_s)
    value = total * args.steps / dev_s

    # ---- end-to-end through the C ABI with host buffers: every step enqueues the pinned H2D of its assignments, the
    # kernels and the D2H of its proof bytes (mp_batch_submit), two steps in flight; N > 1 gathers each step's proofs
    def submit(i):
        nat.check(lib.mp_batch_submit(batches[i], B, z_host.data_ptr(), r_host.data_ptr(), s_host.data_ptr(), out_hosts[i].data_ptr()))

    last_gather = [None]

    def collect(i):
        wait(batches[i])
        if world > 1:
            last_gather[0] = gather_proofs(out_hosts[i].view(B, 192).cuda(non_blocking=True), total, rank, world)

    for i in range(NB):
        out_hosts[i].zero_()
        submit(i)
    for i in range(NB):
        collect(i)
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        i = k % NB
        if k >= NB:
            collect(i)
        submit(i)
    for k in range(max(0, args.steps - NB), args.steps):
        collect(k % NB)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = total * args.steps / e2e_s
    if world > 1:
        gathered = bytes(last_gather[0].cpu().numpy().tobytes()) if rank == 0 else None
    else:
        gathered = bytes(out_hosts[(args.steps - 1) % NB].numpy())

    # ---- single proof (BASELINE configs[1]) on the same context: capacity-1 batch, host buffers in, 192 bytes out
    single = None
    if not args.no_single:
        sb = ctypes.c_void_p()
        nat.check(lib.mp_batch_create_ex(ctx, 1, 1, ctypes.byref(sb)))
        z1 = ctypes.create_string_buffer(z_list[0], n * 32)
        out1 = ctypes.create_string_buffer(192)
        lat, devl = [], []
        for it in range(5 + 20):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            nat.check(lib.mp_batch_upload(sb, 1, z1, r_host.data_ptr(), s_host.data_ptr()))
            ms = ctypes.c_float()
            nat.check(lib.mp_batch_run(sb, ctypes.byref(ms)))
            nat.check(lib.mp_batch_download(sb, out1))
            if it >= 5:
                lat.append((time.perf_counter() - t0) * 1e3)
                devl.append(ms.value)
        buf = (ctypes.c_float * nphase)()
        lib.mp_batch_phase_ms(sb, buf, nphase)
        single = {"e2e_ms_median": statistics.median(lat), "e2e_ms_min": min(lat), "device_ms_median": statistics.median(devl),
                  "phases_ms_last": {lib.mp_phase_name(i).decode(): round(buf[i], 4) for i in range(nphase)},
                  "launches": int(lib.mp_batch_kernel_launches(sb)), "device_bytes": int(lib.mp_batch_device_bytes(sb)),
                  "matches_batch_proof_0": out1.raw == bytes(out_hosts[(args.steps - 1) % NB].numpy()[:192]),
                  "note": "BASELINE configs[1]: upload + kernels + download of ONE proof (mp_batch_upload/run/download on a "
                          "capacity-1 batch = what mp_prove does), wall clock, 20 repetitions after 5 warm-ups"}
        lib.mp_batch_destroy(sb)

    # ---- integer-pipe peak (measured live) and roofline of the dominant kernel
    peak_fq, fqm = int_pipe_peak(nat, lib, local_rank)
    names = [lib.mp_phase_name(i).decode() for i in range(nphase)]
    phase_ms = {names[i]: phase_acc[i] for i in range(nphase)}
    # dominant kernel: round-1 k_ba_bwd<Fq> (first tree level of the four G1 bucket accumulations).  Executed work per affine
    # addition: 4 products (2 back-substitution + lambda + y3) and 1 squaring (lambda^2: 234 of 300 multiplier slots); time
    # from CUDA events recorded on the launching stream around that launch, in the serialised runs above.
    dom_ms, dom_adds = ctypes.c_float(), ctypes.c_uint64()
    nat.check(lib.mp_batch_dominant_kernel(batch, ctypes.byref(dom_ms), ctypes.byref(dom_adds)))
    per_add = 4.0 + SQR_AS_MUL
    achieved = per_add * dom_adds.value / (dom_ms.value * 1e-3) / 1e9 if dom_ms.value > 0 else None
    msm_g1_s = (phase_ms["msm_accumulate_g1"] + phase_ms["msm_reduce_g1"]) * 1e-3
    traffic, ncu = None, {}
    tpath = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(tpath):
        try:
            ncu = json.load(open(tpath))
            # dram bytes per affine addition from the committed `ncu --set full` capture, scaled to this launch
            traffic = ncu["dram_bytes_per_addition"] * dom_adds.value
        except Exception:
            traffic = None
    roofline = {"kernel": "k_ba_bwd<Fq, round 1> (batched-affine bucket trees of the A, B1, L, H MSMs: 1 launch per step, "
                          f"{dom_adds.value} affine additions)", "bound": "integer-pipe",
                "achieved": achieved, "peak": peak_fq / 1e9,
                "unit": "GFq-mul/s (executed: 4 products + 1 squaring = 4.78 product slots per affine addition)",
                "frac": (achieved / (peak_fq / 1e9)) if achieved else None, "traffic": traffic,
                "ms_per_launch": dom_ms.value, "share_of_step": dom_ms.value / (serial_ms / 2) if serial_ms else None,
                "peak_source": "IMAD.WIDE.U32 issue rate measured live (mp_debug_int_pipe_rate) / 300 per 381-bit Montgomery product",
                "measured_fq_mul_rate": fqm / 1e9,
                "ncu_pipe_fmaheavy_pct": ncu.get("sm__pipe_fmaheavy_cycles_active_pct"),
                "ncu_capture": ncu.get("capture"),
                # SURVEY.md 8d credit (pairs x 20 reference windows x 11) over the whole G1 MSM phases / the whole proof: an
                # ALGORITHMIC-savings figure (16 instead of 20 windows, 6.2 instead of 11 products per addition), not a
                # utilisation: it exceeds 1 when the algorithm needs fewer multiplications than the reference's
                "credited_g1_msm": {"gfqmul_per_s": B * CREDIT_G1_PER_PROOF / msm_g1_s / 1e9 if msm_g1_s > 0 else None,
                                    "frac": B * CREDIT_G1_PER_PROOF / msm_g1_s / peak_fq if msm_g1_s > 0 else None},
                "whole_proof": {"credited_gfqmul_per_s": B * args.steps * CREDIT_PER_PROOF / (dev_ms * 1e-3) / 1e9,
                                "frac": B * args.steps * CREDIT_PER_PROOF / (dev_ms * 1e-3) / peak_fq}}
    peaks = measured_peaks()
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    ntt_s = phase_ms["witness_map(r1cs+ntt)"] * 1e-3
    roofline_ntt = {"kernel": "k_ntt_cols + k_ntt_rows (7 transforms of m points per proof)", "bound": "hbm",
                    "achieved": B * NTT_BYTES_PER_PROOF / ntt_s / 1e9 if ntt_s > 0 else None, "peak": hbm_peak, "unit": "GB/s",
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"}
    if roofline_ntt["achieved"]:
        roofline_ntt["frac"] = roofline_ntt["achieved"] / hbm_peak

    # ---- parity of the gathered step (every N) + CPU baseline (rank 0, N = 1 only)
    cpu_baseline, parity = None, None
    if not args.no_cpu_baseline:
        parity, tm = check_parity(cs, pk, rank, world, total, my_idx, z_list, rs, ss, gathered, args.parity_sample,
                                  torch.device("cuda", local_rank))
        if rank == 0 and world == 1 and tm and "all_threads_s_per_proof" in tm:
            cpu_baseline = {"value": 1.0 / tm["all_threads_s_per_proof"], "unit": "proofs/s", "cores": tm["threads"], "kind": "port",
                            "sample": f"{tm['proofs']} {LABEL} proofs of this batch, one at a time on {tm['threads']} host threads (median); 1 more single-threaded",
                            "single_thread_value": 1.0 / tm["single_thread_s_per_proof"] if "single_thread_s_per_proof" in tm else None,
                            "note": "C++ restatement of the reference's arkworks 0.3 path; the reference as shipped is single-threaded"}

    if rank == 0:
        line = {
            "metric": f"{LABEL} Groth16 proofs/sec", "value": value, "unit": "proofs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": DTYPE,
            "data": "synthetic",
            "config": {"workload": f"{WORKLOAD} BLS12-381, batch {B} proofs/GPU/step (BASELINE configs[1] shape, "
                                   f"configs[3] batching)" if SHAPE == "private_transfer" else f"{WORKLOAD} BLS12-381, batch {B} proofs/GPU/step",
                       "proofs_per_step": total,
                       "l2": "inputs larger than L2: 0.37 GB of window tables + the per-batch tree levels and round scratch (GBs) are streamed every step"},
            "e2e": {"value": e2e_value, "unit": "proofs/s", "h2d_bytes_per_step": B * (n * 32 + 64), "d2h_bytes_per_step": B * 192,
                    "ms_per_step": 1e3 * e2e_s / args.steps},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "roofline_ntt": roofline_ntt,
            "cpu_baseline": cpu_baseline, "parity": parity, "single_proof": single,
            "device_bytes": {"per_batch_object": int(lib.mp_batch_device_bytes(batch)), "per_proof": int(lib.mp_batch_device_bytes(batch)) // B,
                             "batch_objects": NB},
            "phase_ms_per_step_serialised": phase_ms, "serialised_ms_per_step": serial_ms / 2,
            "overlap": f"{NB} batch(es) in flight (mp_batch_run_async / mp_batch_submit), batches chain their throughput kernels and overlap copies and latency-bound tails; G2 reduction tail on {'the main' if args.no_g2_stream else 'a second'} stream; phases timed on one stream", "wall_ms_per_step": 1e3 * wall_s / args.steps,
            "setup_s": setup_s,
        }
        print(json.dumps(line), flush=True)
    for bh in batches:
        lib.mp_batch_destroy(bh)
    ctx_obj.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 1 if (parity or "").startswith("MISMATCH") else 0


# ---- single-proof latency (BASELINE configs[1]) -----------------------------------------------------------------------
def run_single(args):
    rank, world, local_rank = env_ranks()
    import manta_rs_b200  # noqa: F401
    from manta_rs_b200 import workload as wl
    cs = wl.make_shape(SHAPE, dist=DIST)
    seeds = [rank * 1000 + i for i in range(4)]
    z_list = make_assignments(cs, seeds, pool=False)
    rs, ss = randomness(seeds, cs.modulus)
    torch, dist = init_cuda(local_rank, world)
    from manta_rs_b200 import _native as nat, keygen, groth16 as g16
    lib = nat.lib()
    g16.Groth16.device = local_rank
    pk = keygen.generate(cs, wl.sample_trapdoor(KEY_SEED), device=local_rank)
    ctx_obj = g16.ProvingContext.decode(pk)
    ctx = ctx_obj.native(g16.R1CS.from_workload(cs, [1] + [0] * (cs.n - 1)).matrices, local_rank)
    n = cs.n
    zb = [ctypes.create_string_buffer(z, n * 32) for z in z_list]
    rb, sb_ = [nat.pack_scalars([r]) for r in rs], [nat.pack_scalars([s]) for s in ss]
    out = ctypes.create_string_buffer(192)
    proofs = {}
    sampler = ClockSampler(local_rank)
    sampler.start()
    for i in range(max(args.warmup, 3)):
        nat.check(lib.mp_prove(ctx, zb[i % 4], rb[i % 4], sb_[i % 4], out))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.begin()
    lat = []
    t_all = time.perf_counter()
    for k in range(args.steps):
        t0 = time.perf_counter()
        nat.check(lib.mp_prove(ctx, zb[k % 4], rb[k % 4], sb_[k % 4], out))
        lat.append((time.perf_counter() - t0) * 1e3)
        proofs[k % 4] = out.raw
    torch.cuda.synchronize()
    total_s = max_over_ranks(time.perf_counter() - t_all)
    clocks = sampler.stop()
    # device-side phases from an explicit capacity-1 batch
    sb = ctypes.c_void_p()
    nat.check(lib.mp_batch_create_ex(ctx, 1, 1, ctypes.byref(sb)))
    nphase = 8
    devl, buf = [], (ctypes.c_float * nphase)()
    for it in range(8):
        nat.check(lib.mp_batch_upload(sb, 1, zb[0], rb[0], sb_[0]))
        ms = ctypes.c_float()
        nat.check(lib.mp_batch_run(sb, ctypes.byref(ms)))
        devl.append(ms.value)
    lib.mp_batch_phase_ms(sb, buf, nphase)
    launches = int(lib.mp_batch_kernel_launches(sb))
    dbytes = int(lib.mp_batch_device_bytes(sb))
    nat.check(lib.mp_batch_set_overlap(sb, 0))
    nat.check(lib.mp_batch_run(sb, None))
    buf2 = (ctypes.c_float * nphase)()
    lib.mp_batch_phase_ms(sb, buf2, nphase)
    lib.mp_batch_destroy(sb)
    peak_fq, fqm = int_pipe_peak(nat, lib, local_rank)
    parity, cpu_baseline = None, None
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import cref, trapdoor
        _, _, trap = trapdoor.key_scalars(cs, wl.sample_trapdoor(KEY_SEED))
        chk = trapdoor.TrapdoorChecker(cs, trap)
        op = cref.OracleProver(pk, cs.p, cs.w, cs.a, cs.b, cs.c)
        threads = host_threads()
        ok, t_cpu = True, []
        for i, pr in proofs.items():
            ok = ok and pr == chk.proof_bytes(unpack_fr(z_list[i]), rs[i], ss[i])
            t0 = time.perf_counter()
            ok = ok and pr == op.prove(z_list[i], rs[i], ss[i], threads=threads)
            t_cpu.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        op.prove(z_list[0], rs[0], ss[0], threads=1)
        t1 = time.perf_counter() - t0
        parity = (f"bit-exact: {len(proofs)} proofs vs the full CPU oracle and the known-trapdoor closed form" if ok else "MISMATCH vs CPU oracle")
        cpu_baseline = {"value": 1.0 / statistics.median(t_cpu), "unit": "proofs/s", "cores": threads, "kind": "port",
                        "sample": f"{len(t_cpu)} {LABEL} proofs on {threads} host threads (median); 1 more single-threaded",
                        "single_thread_value": 1.0 / t1}
    if rank == 0:
        med = statistics.median(lat)
        credit_rate = CREDIT_PER_PROOF / (statistics.median(devl) * 1e-3)
        line = {"metric": f"{LABEL} Groth16 proofs/sec", "value": world * args.steps / total_s, "unit": "proofs/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": med, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": DTYPE, "data": "synthetic",
                "config": {"workload": f"{WORKLOAD} BLS12-381, ONE proof per step through mp_prove with host buffers (BASELINE configs[1])",
                           "l2": "0.37 GB of window tables (> L2) are gathered from every step"},
                "e2e": {"value": 1e3 / med, "unit": "proofs/s", "h2d_bytes_per_step": n * 32 + 64, "d2h_bytes_per_step": 192, "ms_per_step": med,
                        "ms_min": min(lat)},
                "latency_ms": {"median": med, "min": min(lat), "p90": sorted(lat)[int(0.9 * (len(lat) - 1))], "device_median": statistics.median(devl)},
                "phases_ms": {lib.mp_phase_name(i).decode(): round(buf[i], 4) for i in range(nphase)},
                "phases_ms_serialised": {lib.mp_phase_name(i).decode(): round(buf2[i], 4) for i in range(nphase)},
                "gpu_launches": launches * args.steps, "clocks": clocks, "device_bytes": {"per_proof": dbytes},
                "roofline": {"kernel": "whole proof (latency-bound at one proof per step)", "bound": "integer-pipe", "achieved": credit_rate / 1e9,
                             "peak": peak_fq / 1e9, "unit": "GFq-mul/s credited (SURVEY.md 8d)", "frac": credit_rate / peak_fq, "traffic": None},
                "cpu_baseline": cpu_baseline, "parity": parity}
        print(json.dumps(line), flush=True)
    ctx_obj.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 1 if (parity or "").startswith("MISMATCH") else 0


# ---- stand-alone MSM sweep (BASELINE configs[2]) and the sharded G2 stress (configs[4]) ------------------------------
def make_bases(nat, lib, device, group, n, seed):
    """n bases k_i * G generated by the fixed-base kernel from seeded uniform k_i; returns (ark bytes, packed k)."""
    ks = random_scalars(n, seed)
    pb = 96 * group
    out = ctypes.create_string_buffer(n * pb)
    fn = lib.mp_fixed_base_g1 if group == 1 else lib.mp_fixed_base_g2
    nat.check(fn(device, ks, n, out))
    return out, ks


def run_msm_sweep(args):
    rank, world, local_rank = env_ranks()
    torch, dist = init_cuda(local_rank, world)
    import manta_rs_b200  # noqa: F401
    from manta_rs_b200 import _native as nat
    lib = nat.lib()
    peak_fq, fqm = int_pipe_peak(nat, lib, local_rank)
    sizes = []
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches_total = 0
    for log_n in range(args.min_log, args.max_log + 1, 2):
        n = 1 << log_n
        t0 = time.perf_counter()
        bases, ks = make_bases(nat, lib, local_rank, 1, n, 1000 + log_n + 7 * rank)
        sc = random_scalars(n, 2000 + log_n + 7 * rank)
        prep_s = time.perf_counter() - t0
        h = ctypes.c_void_p()
        t0 = time.perf_counter()
        nat.check(lib.mp_msm_bases_create(local_rank, 1, bases, n, ctypes.byref(h)))
        create_s = time.perf_counter() - t0
        sc_pin = torch.frombuffer(bytearray(sc), dtype=torch.uint8).pin_memory()
        out = ctypes.create_string_buffer(96)
        dev, wall = [], []
        for it in range(args.warmup + args.steps):
            ms = ctypes.c_float()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            nat.check(lib.mp_msm_bases_run(h, sc_pin.data_ptr(), n, out, ctypes.byref(ms)))
            if it >= args.warmup:
                wall.append((time.perf_counter() - t0) * 1e3)
                dev.append(ms.value)
        lib.mp_msm_bases_destroy(h)
        credit = credited_msm_fq_muls(n)
        row = {"log_n": log_n, "device_ms": statistics.median(dev), "device_ms_min": min(dev), "e2e_ms": statistics.median(wall),
               "credited_gfqmul_per_s": credit / (statistics.median(dev) * 1e-3) / 1e9, "credited_frac_of_peak": credit / (statistics.median(dev) * 1e-3) / peak_fq,
               "h2d_bytes": n * 32, "prep_s": round(prep_s, 2), "resident_setup_s": round(create_s, 3)}
        if rank == 0 and not args.no_cpu_baseline:
            from oracle import cref
            threads = host_threads()
            if log_n <= args.cpu_max_log:
                ref = ctypes.create_string_buffer(96)
                t0 = time.perf_counter()
                cref.lib().oracle_msm_g1(bases, sc, n, ref, threads)
                row["cpu_oracle_ms"] = (time.perf_counter() - t0) * 1e3
                row["cpu_threads"] = threads
                row["parity"] = "bit-exact vs CPU oracle (full Pippenger)" if ref.raw == out.raw else "MISMATCH"
            else:
                tot = dot_mod(ks, sc, FR)
                row["parity"] = "bit-exact vs closed form (sum k_i s_i) G" if cref.fixed_base(1, [tot]) == out.raw else "MISMATCH"
        sizes.append(row)
        del bases
    clocks = sampler.stop()
    if world > 1:
        dist.barrier()
    if rank == 0:
        top = sizes[-1]
        bad = any(r.get("parity", "").startswith("MISMATCH") for r in sizes)
        cpu = [r for r in sizes if "cpu_oracle_ms" in r]
        cpu_baseline = None
        if cpu:
            c = cpu[-1]
            cpu_baseline = {"value": credited_msm_fq_muls(1 << c["log_n"]) / (c["cpu_oracle_ms"] * 1e-3) / 1e9, "unit": "GFq-mul/s", "cores": c["cpu_threads"],
                            "kind": "port", "sample": f"one 2^{c['log_n']}-point G1 MSM (ark window rule, Jacobian buckets) on {c['cpu_threads']} host threads"}
        line = {"metric": "G1 MSM credited Fq-mul/s vs integer-pipe roofline", "value": top["credited_gfqmul_per_s"] * world, "unit": "GFq-mul/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": top["device_ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": DTYPE, "data": "synthetic",
                "config": {"workload": f"stand-alone BLS12-381 G1 variable-base MSM sweep 2^{args.min_log}..2^{args.max_log} (BASELINE configs[2]); bases k_i G resident "
                                       f"(mp_msm_bases_*), uniform Fr scalars from pinned host memory; value = the 2^{top['log_n']} point; replicas only at N > 1",
                           "l2": "bases + window rows exceed L2 from 2^18 on; below that the step is latency-bound"},
                "e2e": {"value": credited_msm_fq_muls(1 << top["log_n"]) / (top["e2e_ms"] * 1e-3) / 1e9, "unit": "GFq-mul/s", "h2d_bytes_per_step": top["h2d_bytes"],
                        "d2h_bytes_per_step": 96, "ms_per_step": top["e2e_ms"]},
                "sizes": sizes, "clocks": clocks, "gpu_launches": None,
                "roofline": {"kernel": f"whole MSM at 2^{top['log_n']} (sort + bucket trees + reduction)", "bound": "integer-pipe", "achieved": top["credited_gfqmul_per_s"],
                             "peak": peak_fq / 1e9, "unit": "GFq-mul/s credited (N x ceil(255/c_ref) x 11, SURVEY.md 8d)", "frac": top["credited_frac_of_peak"], "traffic": None},
                "cpu_baseline": cpu_baseline, "parity": "MISMATCH" if bad else "every size bit-exact (oracle Pippenger up to 2^%d, closed form above)" % args.cpu_max_log}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_g2_stress(args):
    rank, world, local_rank = env_ranks()
    torch, dist = init_cuda(local_rank, world)
    import manta_rs_b200  # noqa: F401
    from manta_rs_b200 import _native as nat, sharded
    lib = nat.lib()
    group, pb = args.group, 96 * args.group
    n = 1 << args.log_n
    lo, hi = sharded.shard_range(n, rank, world)
    cnt = hi - lo
    # every rank builds only its own slice of the (seeded, globally defined) bases and scalars
    t0 = time.perf_counter()
    ks_all = random_scalars(n, 4242)
    sc_all = random_scalars(n, 2424)
    ks, sc = ks_all[32 * lo:32 * hi], sc_all[32 * lo:32 * hi]
    bases = ctypes.create_string_buffer(max(cnt, 1) * pb)
    fn = lib.mp_fixed_base_g1 if group == 1 else lib.mp_fixed_base_g2
    nat.check(fn(local_rank, ks, cnt, bases))
    prep_s = time.perf_counter() - t0
    h = ctypes.c_void_p()
    nat.check(lib.mp_msm_bases_create(local_rank, group, bases, cnt, ctypes.byref(h)))
    sc_pin = torch.frombuffer(bytearray(sc), dtype=torch.uint8).pin_memory()
    dev = torch.device("cuda", local_rank)
    part = ctypes.create_string_buffer(pb)
    sum_fn = lib.mp_points_sum_g1 if group == 1 else lib.mp_points_sum_g2
    peak_fq, fqm = int_pipe_peak(nat, lib, local_rank)

    def step():
        ms = ctypes.c_float()
        nat.check(lib.mp_msm_bases_run(h, sc_pin.data_ptr(), cnt, part, ctypes.byref(ms)))
        if world == 1:
            return part.raw, ms.value
        mine = torch.frombuffer(bytearray(part.raw), dtype=torch.uint8).to(dev)
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        stacked = b"".join(bytes(p.cpu().numpy()) for p in parts)
        out = ctypes.create_string_buffer(pb)
        nat.check(sum_fn(local_rank, stacked, world, out))
        return out.raw, ms.value

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.begin()
    devs, result = [], None
    t0 = time.perf_counter()
    for _ in range(args.steps):
        result, ms = step()
        devs.append(ms)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    dev_ms = max_over_ranks(statistics.median(devs))
    clocks = sampler.stop()
    lib.mp_msm_bases_destroy(h)
    same = gather_by_rank(result, rank, world, dev)
    if rank == 0:
        parity, cpu_baseline = None, None
        if not args.no_cpu_baseline:
            from oracle import cref
            tot = dot_mod(ks_all, sc_all, FR)
            ok = cref.fixed_base(group, [tot]) == result and all(x == result for x in same)
            parity = (f"bit-exact: result on all {world} rank(s) == (sum k_i s_i) G{group} from the CPU oracle's fixed-base routine" if ok else "MISMATCH")
            # bounded CPU sample: the first 2^14 (G2) / 2^16 (G1) pairs of rank 0's slice through the oracle's Pippenger
            sn = min(cnt, 1 << (14 if group == 2 else 16))
            ref = ctypes.create_string_buffer(pb)
            threads = host_threads()
            t1 = time.perf_counter()
            (cref.lib().oracle_msm_g1 if group == 1 else cref.lib().oracle_msm_g2)(bases, sc, sn, ref, threads)
            dt = time.perf_counter() - t1
            cpu_baseline = {"value": credited_msm_fq_muls(sn, group == 2) / dt / 1e9, "unit": "GFq-mul/s", "cores": threads, "kind": "port",
                            "sample": f"one 2^{sn.bit_length() - 1}-point G{group} MSM on {threads} host threads ({dt * 1e3:.0f} ms)"}
        credit = credited_msm_fq_muls(n, group == 2)
        line = {"metric": f"G{group} MSM credited Fq-mul/s (one 2^{args.log_n}-base MSM sharded by base range)", "value": credit / (dev_ms * 1e-3) / 1e9,
                "unit": "GFq-mul/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
                "config": {"workload": f"ONE BLS12-381 G{group} MSM over 2^{args.log_n} bases (trusted-setup-sized: kzg.rs:43-44, 509-523; BASELINE configs[4]), "
                                       f"base range sharded over {world} rank(s), partial sums combined by one {pb}-byte all_gather + mp_points_sum",
                           "l2": "bases + window rows of a slice exceed L2"},
                "e2e": {"value": credit / (e2e_s / args.steps) / 1e9, "unit": "GFq-mul/s", "h2d_bytes_per_step": cnt * 32, "d2h_bytes_per_step": pb,
                        "ms_per_step": 1e3 * e2e_s / args.steps},
                "clocks": clocks, "gpu_launches": None, "exchange_bytes_per_rank": pb, "prep_s": round(prep_s, 2),
                "roofline": {"kernel": "whole sharded MSM (sort + bucket trees + reduction of the largest slice)", "bound": "integer-pipe",
                             "achieved": credit / (dev_ms * 1e-3) / 1e9, "peak": world * peak_fq / 1e9,
                             "unit": "GFq-mul/s credited (N x ceil(255/c_ref) x 33 in G2, SURVEY.md 8d)", "frac": credit / (dev_ms * 1e-3) / (world * peak_fq), "traffic": None},
                "cpu_baseline": cpu_baseline, "parity": parity}
        print(json.dumps(line), flush=True)
        rc = 1 if (parity or "").startswith("MISMATCH") else 0
    else:
        rc = 0
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return rc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=int(os.environ.get("MP_BENCH_BATCH", "128")), help="proofs per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="prove", choices=["prove", "single", "msm_sweep", "g2_stress"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU oracle legs (parity check and cpu_baseline)")
    ap.add_argument("--parity-sample", type=int, default=32, help="proofs of the last step checked against the full CPU oracle")
    ap.add_argument("--no-single", action="store_true", help="prove workload: skip the single-proof latency sub-measurement")
    ap.add_argument("--shape", default="private_transfer", choices=sorted(SHAPE_INFO),
                    help="circuit shape (default: the headline PrivateTransfer workload)")
    ap.add_argument("--dist", default="U", choices=["U", "R"], help="witness distribution (SURVEY.md 8d config 2)")
    ap.add_argument("--inflight", type=int, default=int(os.environ.get("MP_BENCH_INFLIGHT", "2")), choices=[1, 2],
                    help="batches in flight per GPU (tuning knob; default 2)")
    ap.add_argument("--no-g2-stream", action="store_true", help="run the G2 MSM on the main stream (tuning knob)")
    ap.add_argument("--min-log", type=int, default=16)
    ap.add_argument("--max-log", type=int, default=24, help="msm_sweep: largest size (2^max_log points)")
    ap.add_argument("--cpu-max-log", type=int, default=20, help="msm_sweep: largest size the CPU oracle's Pippenger is run on")
    ap.add_argument("--log-n", type=int, default=20, help="g2_stress: bases = 2^log_n")
    ap.add_argument("--group", type=int, default=2, choices=[1, 2], help="g2_stress: group of the sharded MSM")
    args = ap.parse_args()
    set_shape(args.shape, args.dist)
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)
    return {"prove": run_prove, "single": run_single, "msm_sweep": run_msm_sweep, "g2_stress": run_g2_stress}[args.workload](args)


if __name__ == "__main__":
    sys.exit(main())
