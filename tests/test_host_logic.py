"""Host-side logic that needs no GPU: rng mirror, workload generator, codecs, the C ABI surface."""
import ctypes
import os
import re
import subprocess

import pytest

from helpers import ROOT, BLS12_381 as C
import manta_rs_b200.workload as wl
from manta_rs_b200 import rng as mrng


def test_chacha20_known_answer():
    # djb ChaCha20, all-zero key and nonce, first 64 bytes of key stream (the vector rand_chacha 0.3 reproduces)
    r = mrng.ChaCha20Rng(bytes(32))
    assert r.fill_bytes(64).hex() == (
        "76b8e0ada0f13d90405d6ae55386bd28bdd219b8a08ded1aa836efcc8b770dc7"
        "da41597c5157488d7724e03fb8d84a376a43b8f41518a11cc387b669b2ee6586")
    # next_u64 = lo | hi << 32 of consecutive words; straddles buffer refills transparently
    a, b = mrng.ChaCha20Rng(bytes(range(32))), mrng.ChaCha20Rng(bytes(range(32)))
    for _ in range(100):
        lo, hi = b.next_u32(), b.next_u32()
        assert a.next_u64() == lo | (hi << 32)


def test_field_rand_rule():
    """ark-ff 0.3 `Fr::rand`: limbs from next_u64, shave, reject >= modulus, limbs ARE the Montgomery form."""
    class Fixed:
        def __init__(self, vals):
            self.vals = list(vals)

        def next_u64(self):
            return self.vals.pop(0)

    R = 1 << 256
    limbs = [5, 6, 7, (1 << 63) | 9]                       # top bit is shaved off (REPR_SHAVE_BITS = 1)
    v = 5 | (6 << 64) | (7 << 128) | (9 << 192)
    assert mrng.field_rand(Fixed(limbs), C.r) == v * pow(R, -1, C.r) % C.r
    # a draw >= modulus is rejected and redrawn
    big = [0xFFFFFFFFFFFFFFFF] * 4
    assert mrng.field_rand(Fixed(big + limbs), C.r) == v * pow(R, -1, C.r) % C.r


def test_workload_shapes_and_satisfaction():
    for name, (n, p, w, log_m) in wl.SHAPES.items():
        assert n == p + w and (1 << (log_m - 1)) < w + p <= (1 << log_m)
    cs = wl.make_r1cs(4, 200, dist="R")
    z = wl.make_assignment(cs, 7)
    assert z[0] == 1 and wl.is_satisfied(cs, z)
    assert wl.make_assignment(cs, 7) == z and wl.make_assignment(cs, 8) != z
    kinds = set(cs.kinds)
    assert wl.KIND_BOOL in kinds and wl.KIND_MUL in kinds
    for i, k in enumerate(cs.kinds):
        if k == wl.KIND_BOOL:
            assert z[cs.p + i] in (0, 1)
        if k == wl.KIND_SMALL:
            assert z[cs.p + i] < (1 << 128)
    z[-1] = (z[-1] + 1) % C.r
    assert not wl.is_satisfied(cs, z)


def test_header_symbols_exported_and_bound(native):
    """Every function include/mantaprover.h declares is exported by the built library and bound in _native.py."""
    hdr = open(os.path.join(ROOT, "include", "mantaprover.h")).read()
    declared = set(re.findall(r"MP_API [\w\s\*]+?\b(mp_\w+)\(", hdr))
    assert len(declared) >= 25
    out = subprocess.run(["nm", "-D", "--defined-only", native.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (mp_\w+)", out))
    assert declared <= exported, declared - exported
    assert declared == set(native.SYMBOLS), declared ^ set(native.SYMBOLS)
    lib = native.lib()
    assert lib.mp_strerror(0) == b"ok" and b"fallback" in lib.mp_strerror(3)


def test_pk_parse_and_errors_without_gpu(native):
    from oracle.pyref import groth16 as og
    from manta_rs_b200 import groth16 as g16
    cs = wl.make_r1cs(2, 6)
    pk, _ = og.setup_trapdoor(C, cs.as_dict(), *wl.sample_trapdoor(1))
    pkb = og.pk_to_bytes(C, pk)
    view = native.PkView()
    buf = ctypes.create_string_buffer(pkb, len(pkb))
    assert native.lib().mp_pk_parse(buf, len(pkb), ctypes.byref(view)) == 0
    assert (view.a_len, view.b_g1_len, view.b_g2_len, view.h_len, view.l_len, view.gamma_abc_len) == (8, 8, 8, 7, 6, 2)
    base = ctypes.addressof(buf)
    assert view.alpha_g1 == base and view.beta_g2 == base + 96
    # truncated / trailing bytes are format errors, never crashes
    assert native.lib().mp_pk_parse(buf, len(pkb) - 1, ctypes.byref(view)) == 5
    with pytest.raises(g16.Error):
        g16.ProvingContext.decode(pkb + b"\0")
    ctx = g16.ProvingContext.decode(pkb)
    assert ctx.encode() == pkb and ctx == ctx.clone()
    # Proof codec: `codec::Encode` wraps the 192 bytes as Vec<u8> (u64-LE length prefix)
    pr = g16.Proof(bytes(range(192)))
    assert pr.encode() == (192).to_bytes(8, "little") + bytes(range(192))
    with pytest.raises(g16.Error):
        g16.Proof(b"short")


def test_no_cpu_fallback_without_device(native):
    """Without a CUDA device every compute entry point fails loudly (MP_ERR_NO_DEVICE), it never computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present")
    out = ctypes.create_string_buffer(96)
    ms = ctypes.c_float()
    rc = native.lib().mp_msm_g1(0, None, None, 0, out, ctypes.byref(ms))
    assert rc in (2, 3)
    with pytest.raises(native.NativeError):
        native.check(rc)
    data = ctypes.create_string_buffer(64)
    assert native.lib().mp_ntt(0, data, 1, 0, 0, None) in (2, 3)


def test_chacha20_block_rfc7539_vector():
    """The ChaCha20 block function of the rng mirror against RFC 7539 section 2.3.2 (key 00..1f, block counter 1, nonce
    00:00:00:09:00:00:00:4a:00:00:00:00): in the djb layout rand_chacha uses, state words 12..15 are counter_lo, counter_hi,
    stream_lo, stream_hi, so the RFC's (counter, nonce) is counter = 1 | 0x09000000 << 32, stream = 0x4a000000."""
    import struct
    from manta_rs_b200.rng import chacha20_block, ChaCha20Rng
    key = struct.unpack("<8I", bytes(range(32)))
    out = chacha20_block(key, 1 | (0x09000000 << 32), 0x4A000000)
    want = ("e4e7f110 15593bd1 1fdd0f50 c47120a3 c7f4d1c7 0368c033 9aaa2204 4e6cd4c3 "
            "466482d2 09aa9f07 05d7c214 a2028bd9 d19c12b5 b94e16de e883d0cb 4e3c50a2")
    assert " ".join("%08x" % w for w in out) == want
    # word stream: block 0 then block 1 of the same key, stream 0; next_u64 = lo | hi << 32
    rng = ChaCha20Rng(bytes(range(32)))
    b0 = chacha20_block(key, 0, 0)
    assert rng.next_u64() == b0[0] | b0[1] << 32
    for _ in range(7):
        rng.next_u64()
    b1 = chacha20_block(key, 1, 0)
    assert rng.next_u32() == b1[0]


def test_batched_affine_pair_capacity_covers_worst_case_bucket_loads():
    """Level r of the bucket trees adds floor(ceil(c / 2^(r-1)) / 2) pairs in a bucket with c entries.  The launcher sizes every
    level for a capacity computed from the geometry alone (the bucket loads are only known on the device): it must cover every
    distribution of at most max_entries entries over the buckets, and the provisioned number of levels must reach a bucket
    that holds every entry."""
    import ctypes
    import random
    from manta_rs_b200 import _native as nat
    lib = nat.lib()

    def geometry(c, groups, n, r):
        cap, rounds, nb, me = ctypes.c_uint32(), ctypes.c_int(), ctypes.c_uint32(), ctypes.c_uint32()
        nat.check(lib.mp_debug_ba_geometry(c, groups, n, r, ctypes.byref(cap), ctypes.byref(rounds), ctypes.byref(nb), ctypes.byref(me)))
        return cap.value, rounds.value, nb.value, me.value

    def pairs(counts, r):
        return sum((((c + (1 << (r - 1)) - 1) >> (r - 1)) >> 1) for c in counts)

    rng = random.Random(4)
    for c, groups, n in ((16, 1, 35179), (16, 1, 65536), (13, 0, 70000), (4, 0, 9), (8, 1, 1000), (16, 1, 1)):
        _, rounds, nb, me = geometry(c, groups, n, 1)
        windows = 255 // c + 1
        geff = windows if (groups <= 0 or groups > windows) else groups
        fullest = min(me, n * -(-windows // geff))      # one entry per (scalar, table row) at most
        assert (1 << rounds) >= fullest
        loads = []
        loads.append([me] + [0] * (nb - 1))                                     # everything in one bucket
        base, extra = divmod(me, nb)
        loads.append([base + (1 if i < extra else 0) for i in range(nb)])       # as even as possible
        for r in range(2, 8):                                                   # buckets of 2^(r-1) + 1 entries: one pair each at level r
            k = (1 << (r - 1)) + 1
            full = min(nb, me // k)
            loads.append([k] * full + [0] * (nb - full))
        loads.append([3] * min(nb, me // 3) + [0] * (nb - min(nb, me // 3)))
        for _ in range(3):                                                      # random skew
            left, v = me, []
            for _ in range(nb):
                x = min(left, int(rng.expovariate(nb / max(me, 1)) * 2))
                v.append(x)
                left -= x
            loads.append(v)
        for counts in loads:
            assert sum(counts) <= me and len(counts) == nb
            for r in range(1, rounds + 1):
                cap = geometry(c, groups, n, r)[0]
                assert pairs(counts, r) <= cap, (c, groups, n, r, pairs(counts, r), cap)
            assert pairs(counts, rounds + 1) == 0 or max(counts) > (1 << rounds)
        # the fullest possible bucket is exhausted by the provisioned levels
        assert pairs([fullest] + [0] * (nb - 1), rounds + 1) == 0 and (fullest < 2 or pairs([fullest], rounds) >= 1)


def test_batched_affine_round_scratch_covers_every_partial_batch():
    """A batch object re-plans its rounds for the live count (T pairs per thread grows with the pairs of the launch), so a
    partial batch just under a T threshold needs MORE thread slots than a full one.  The scratch is sized for a slab of the
    capacity (all of it up to 32 vectors, half of it above) and must cover EVERY count up to that slab at the three
    reference shapes (counts 65..126 of 128 on ToPublic, 36..49 of 64 on PrivateTransfer used to overrun it);
    larger live counts run a level in several launches of at most one slab."""
    import ctypes
    from manta_rs_b200 import _native as nat
    lib = nat.lib()
    out = (ctypes.c_uint64 * 5)()
    shapes = {"to_private": (8253, 1 << 14), "private_transfer": (35175, 1 << 16), "to_public": (27945, 1 << 15), "tiny": (64, 64)}
    for name, (n, m) in shapes.items():
        for cap in (1, 2, 16, 64, 128, 228, 512):
            for g2 in (0, 1):
                worst = 0.0
                nat.check(lib.mp_debug_prove_ba_demand(n, m, cap, cap, g2, out))
                slab = out[4]
                assert slab == (cap if cap <= 32 else max(32, -(-cap // 2)))
                for count in range(1, slab + 1):
                    nat.check(lib.mp_debug_prove_ba_demand(n, m, cap, count, g2, out))
                    need_pairs, need_threads, have_pairs, have_threads = out[0], out[1], out[2], out[3]
                    assert need_pairs <= have_pairs and need_threads <= have_threads, (name, cap, count, g2, list(out))
                    worst = max(worst, need_threads / have_threads)
                assert worst == 1.0, (name, cap, g2, worst)   # the bound is tight: some count needs exactly the provisioned slots
    assert lib.mp_debug_prove_ba_demand(35175, 1 << 16, 4, 5, 0, out) == 1


def test_bench_helpers_scalars_credit_and_dot():
    """Host-side helpers of bench.py: uniform canonical scalars, the reference's window rule behind the credited work
    (SURVEY.md 8d figures), and the closed-form dot product."""
    import bench
    r = bench.FR
    buf = bench.random_scalars(5000, 3)
    vals = bench.unpack_fr(buf)
    assert len(vals) == 5000 and all(v < r for v in vals) and max(vals) > r // 2 and len(set(vals)) == 5000
    assert bench.random_scalars(5000, 3) == buf and bench.random_scalars(10, 4) != buf[:320]
    assert [bench.ark_window(1 << k) for k in (16, 18, 20, 22, 24)] == [13, 14, 15, 17, 18] and bench.ark_window(31) == 3
    assert bench.credited_msm_fq_muls(1 << 16) == 14417920 and bench.credited_msm_fq_muls(1 << 20) == 196083712   # 14.42 M / 196.1 M
    assert bench.credited_msm_fq_muls(1 << 24) == 2768240640 and bench.credited_msm_fq_muls(1 << 20, True) == 3 * 196083712
    other = bench.random_scalars(5000, 5)
    assert bench.dot_mod(buf, other, r) == sum(a * b for a, b in zip(vals, bench.unpack_fr(other))) % r


def test_prover_pool_requeues_the_proofs_of_a_failed_device():
    """Fault path: a device that fails is retired and its chunk is proved by the survivors; only when every
    device has failed does the call collapse to the opaque `Error` of groth16.rs:50-60.  CPU stand-ins for the device call."""
    import threading
    import time
    from manta_rs_b200 import groth16 as g16, _native as nat

    class FakeCompiler:
        matrices = object()

        def __init__(self, i):
            self.assignment = [1, i]

    comps = [FakeCompiler(i) for i in range(37)]
    for c in comps:
        c.matrices = FakeCompiler.matrices
    rs, ss = list(range(37)), list(range(100, 137))
    calls, lock = [], threading.Lock()

    def fake(fail_devices):
        def prove_chunk(device, compilers, r, s):
            with lock:
                calls.append((device, [c.assignment[1] for c in compilers]))
            time.sleep(0.02)            # a chunk takes long enough for every worker to pick one up
            if device in fail_devices:
                raise nat.NativeError(2, "CUDA runtime error", "injected")
            return [bytes([c.assignment[1]]) * 192 for c in compilers]
        return prove_chunk

    pool = g16.ProverPool(context=None, devices=[0, 1, 2], chunk=5, prove_chunk=fake({1}))
    out = pool.prove_many_with_randomness(comps, rs, ss)
    assert [p.to_bytes()[0] for p in out] == list(range(37))
    assert list(pool.failed) == [1] and sum(1 for d, _ in calls if d == 1) == 1      # retired after its first failure
    proved = sorted(i for d, idx in calls if d != 1 for i in idx)
    assert proved == list(range(37))                                                 # the failed chunk was re-proved elsewhere
    # the pool keeps the device retired on the next call
    calls.clear()
    pool.prove_many_with_randomness(comps[:6], rs[:6], ss[:6])
    assert all(d != 1 for d, _ in calls)
    with pytest.raises(g16.Error):
        g16.ProverPool(context=None, devices=[0, 1], chunk=4, prove_chunk=fake({0, 1})).prove_many_with_randomness(comps, rs, ss)
    assert g16.ProverPool(context=None, devices=[0], prove_chunk=fake(set())).prove_many_with_randomness([], [], []) == []
