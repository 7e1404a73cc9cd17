"""Parity of the CUDA path against the CPU oracle — bit-exact, through the C ABI.  Run with `-m gpu` on the B200."""
import ctypes
import json
import os
import random

import pytest

from helpers import cref, BLS12_381 as C, oracle_keygen, trapdoor_proof_bytes
import manta_rs_b200.workload as wl

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _chk(native, rc):
    native.check(rc)


# ---- field and group arithmetic -----------------------------------------------------------------------------
@pytest.mark.parametrize("field", [0, 1])
def test_field_ops(native, field):
    p, limbs = ((C.q, 6), (C.r, 4))[field]
    rng = random.Random(field)
    n = 4096
    a = [rng.randrange(p) for _ in range(n)]
    b = [rng.randrange(p) for _ in range(n)]
    a[:4] = [0, p - 1, 1, p - 1]
    b[:4] = [0, p - 1, p - 1, 1]
    for op in range(6):
        out = ctypes.create_string_buffer(n * limbs * 8)
        _chk(native, native.lib().mp_debug_field_op(0, field, op, native.pack_scalars(a, limbs), native.pack_scalars(b, limbs), out, n))
        assert native.unpack_scalars(out.raw, limbs) == cref.field_op(field, op, a, b if op < 3 else None), (field, op)
    # op 6: inversion by the binary extended Euclid (used by the batched-affine MSM) == Fermat inversion of the oracle
    a[4:8] = [2, p - 2, (p + 1) // 2, 1 << 200]
    a[8:12] = [3, p - 3, (1 << 64) - 1, (1 << 63) + 5]
    want = cref.field_op(field, 4, a, None)
    for op in (6, 7):
        out = ctypes.create_string_buffer(n * limbs * 8)
        _chk(native, native.lib().mp_debug_field_op(0, field, op, native.pack_scalars(a, limbs), native.pack_scalars(b, limbs), out, n))
        assert native.unpack_scalars(out.raw, limbs) == want, (field, "inv_gcd", op)


@pytest.mark.parametrize("group", [1, 2])
def test_group_ops(native, group):
    from oracle.pyref.curves import Group
    G = Group(C, group)
    rng = random.Random(group)
    n, pb = 48, 96 * group
    ka = [rng.randrange(1, C.r) for _ in range(n)]
    kb = [rng.randrange(1, C.r) for _ in range(n)]
    kb[0], kb[1] = ka[0], C.r - ka[1]            # doubling and inverse cases
    A = bytearray(cref.fixed_base(group, ka))
    B = bytearray(cref.fixed_base(group, kb))
    inf = G.serialize_uncompressed(None)
    A[2 * pb:3 * pb] = inf
    B[3 * pb:4 * pb] = inf
    A[4 * pb:5 * pb] = inf
    B[4 * pb:5 * pb] = inf
    ka[2] = ka[4] = 0
    kb[3] = kb[4] = 0
    ks = [rng.randrange(C.r) for _ in range(n)]
    ks[5], ks[6], ks[7] = 0, 1, C.r - 1
    out = ctypes.create_string_buffer(n * pb)
    _chk(native, native.lib().mp_debug_group_op(0, group, 0, bytes(A), bytes(B), None, out, n))
    assert out.raw == cref.fixed_base(group, [(x + y) % C.r for x, y in zip(ka, kb)])
    _chk(native, native.lib().mp_debug_group_op(0, group, 1, bytes(A), None, None, out, n))
    assert out.raw == cref.fixed_base(group, [2 * x % C.r for x in ka])
    _chk(native, native.lib().mp_debug_group_op(0, group, 2, bytes(A), None, native.pack_scalars(ks), out, n))
    assert out.raw == cref.fixed_base(group, [x * k % C.r for x, k in zip(ka, ks)])


def test_fixed_base_matches_oracle(native):
    rng = random.Random(11)
    ks = [0, 1, C.r - 1] + [rng.randrange(C.r) for _ in range(200)]
    for group, fn, pb in ((1, native.lib().mp_fixed_base_g1, 96), (2, native.lib().mp_fixed_base_g2, 192)):
        out = ctypes.create_string_buffer(len(ks) * pb)
        _chk(native, fn(0, native.pack_scalars(ks), len(ks), out))
        assert out.raw == cref.fixed_base(group, ks)


# ---- MSM ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("group,n", [(1, 0), (1, 1), (1, 31), (1, 1000), (1, 1 << 14), (2, 0), (2, 1), (2, 600), (2, 1 << 12)])
def test_msm_vs_oracle(native, group, n):
    rng = random.Random(100 * group + n)
    pb = 96 * group
    ks = [rng.randrange(1, C.r) for _ in range(n)]
    bases = bytearray(cref.fixed_base(group, ks))
    sc = [rng.randrange(C.r) for _ in range(n)]
    if n >= 31:
        sc[0], sc[1], sc[2], sc[3], sc[4] = 0, 1, C.r - 1, (1 << 128) - 1, 1 << 254
        from oracle.pyref.curves import Group
        bases[7 * pb:8 * pb] = Group(C, group).serialize_uncompressed(None)       # infinity base
        bases[9 * pb:10 * pb] = bases[8 * pb:9 * pb]                               # repeated base, same scalar
        sc[9] = sc[8]
        for i in range(10, 20):                                                    # many zeros / ones (ark shortcuts)
            sc[i] = i & 1
    out = ctypes.create_string_buffer(pb)
    ms = ctypes.c_float()
    fn = native.lib().mp_msm_g1 if group == 1 else native.lib().mp_msm_g2
    _chk(native, fn(0, bytes(bases), native.pack_scalars(sc), n, out, ctypes.byref(ms)))
    assert out.raw == cref.msm(group, bytes(bases), sc, threads=8)


def test_msm_skewed_scalars(native):
    """All scalars equal (one giant bucket per window) exercises the segment splitting."""
    n = 5000
    rng = random.Random(5)
    bases = cref.fixed_base(1, [rng.randrange(1, C.r) for _ in range(n)])
    for val in (1, 3, (1 << 255) % C.r, C.r - 1):
        sc = [val] * n
        out = ctypes.create_string_buffer(96)
        _chk(native, native.lib().mp_msm_g1(0, bases, native.pack_scalars(sc), n, out, None))
        assert out.raw == cref.msm(1, bases, sc, threads=8)


@pytest.mark.parametrize("group", [1, 2])
def test_msm_degenerate_pairs(native, group):
    """Equal and opposite points inside one bucket: the pairwise trees of the batched-affine accumulation meet doublings
    (P + P, then 2P + 2P), cancellations (P + (-P)), infinities as operands, and an MSM whose value is the point at infinity."""
    from oracle.pyref.curves import Group
    G = Group(C, group)
    rng = random.Random(77 + group)
    pb = 96 * group
    k = rng.randrange(1, C.r)
    P = cref.fixed_base(group, [k])
    negP = cref.fixed_base(group, [C.r - k])
    inf = G.serialize_uncompressed(None)
    other = cref.fixed_base(group, [rng.randrange(1, C.r) for _ in range(40)])
    fn = native.lib().mp_msm_g1 if group == 1 else native.lib().mp_msm_g2
    cases = []
    s0 = rng.randrange(C.r)
    cases.append((P * 8, [s0] * 8))                                   # 8 copies: three levels of doublings
    cases.append((P * 5 + negP * 5, [s0] * 10))                       # cancels to infinity
    cases.append((P + negP + P + inf + negP + P, [s0] * 6))           # mixed, result s0 * P
    cases.append((other + P * 3 + negP * 2 + inf * 3, [rng.randrange(C.r) for _ in range(40)] + [s0] * 8))
    cases.append((P + negP, [s0, s0]))
    cases.append((inf * 7, [s0] * 7))
    for bases, sc in cases:
        n = len(sc)
        assert len(bases) == n * pb
        out = ctypes.create_string_buffer(pb)
        _chk(native, fn(0, bytes(bases), native.pack_scalars(sc), n, out, None))
        assert out.raw == cref.msm(group, bytes(bases), sc, threads=1)


@pytest.mark.parametrize("group", [1, 2])
def test_points_sum_and_sharded_msm(native, group):
    """An MSM split by base range into 3 slices (the multi-GPU sharding of BASELINE configs[4], here on one device):
    partial sums + mp_points_sum == the single MSM == the oracle."""
    from manta_rs_b200 import sharded
    rng = random.Random(31 + group)
    n, pb = 700, 96 * group
    bases = cref.fixed_base(group, [rng.randrange(1, C.r) for _ in range(n)])
    sc = [rng.randrange(C.r) for _ in range(n)]
    scalars = native.pack_scalars(sc)
    parts = b""
    for r in range(3):
        lo, hi = sharded.shard_range(n, r, 3)
        part, _ = sharded._native_msm(group, 0)(bases[lo * pb:hi * pb], scalars[lo * 32:hi * 32], hi - lo)
        parts += part
    total = sharded._native_sum(group, 0)(parts, 3)
    whole, _ = sharded.msm_sharded(group, bases, scalars, rank=0, world=1)
    assert total == whole == cref.msm(group, bases, sc, threads=8)
    from oracle.pyref.curves import Group
    inf = Group(C, group).serialize_uncompressed(None)
    assert sharded._native_sum(group, 0)(b"", 0) == inf
    assert sharded._native_sum(group, 0)(inf + parts[:pb] + inf, 3) == parts[:pb]


@pytest.mark.parametrize("group,n", [(1, 3000), (2, 700), (1, 5)])
def test_msm_resident_bases(native, group, n):
    """Bases resident on the device (mp_msm_bases_*): several scalar vectors against one handle, fewer scalars than bases
    (ark: size = min(bases, scalars)), and an empty call; with and without room for the precomputed window rows."""
    lib = native.lib()
    rng = random.Random(50 + group + n)
    pb = 96 * group
    bases = cref.fixed_base(group, [rng.randrange(1, C.r) for _ in range(n)])
    for limit in (None, "0"):
        if limit is not None:
            os.environ["MP_MSM_TABLE_LIMIT_MB"] = limit     # no precomputed rows: one bucket set per window + Horner
        try:
            h = ctypes.c_void_p()
            _chk(native, lib.mp_msm_bases_create(0, group, bases, n, ctypes.byref(h)))
        finally:
            os.environ.pop("MP_MSM_TABLE_LIMIT_MB", None)
        for cnt in (n, n, n // 2, 1, 0):
            sc = [rng.randrange(C.r) for _ in range(cnt)]
            out = ctypes.create_string_buffer(pb)
            ms = ctypes.c_float()
            _chk(native, lib.mp_msm_bases_run(h, native.pack_scalars(sc), cnt, out, ctypes.byref(ms)))
            assert out.raw == cref.msm(group, bases[:cnt * pb], sc, threads=8), (group, n, cnt, limit)
        assert lib.mp_msm_bases_run(h, native.pack_scalars([1] * (n + 1)), n + 1, out, None) == 1   # more scalars than bases
        lib.mp_msm_bases_destroy(h)


def test_non_canonical_scalars_are_rejected(native):
    """z, r, s cross the ABI canonical (< r); a value >= r fails the call instead of yielding a wrong proof."""
    from manta_rs_b200 import groth16 as g16
    cs = wl.make_r1cs(3, 40)
    pk, trap = oracle_keygen(cs, wl.sample_trapdoor(13))
    ctx = g16.ProvingContext.decode(pk)
    z = wl.make_assignment(cs, 1)
    h = ctx.native(g16.R1CS.from_workload(cs, z).matrices)
    out = ctypes.create_string_buffer(192)
    lib = native.lib()
    raw = lambda vals: b"".join(int(v).to_bytes(32, "little") for v in vals)
    assert lib.mp_prove(h, raw(z), raw([5]), raw([6]), out) == 0
    assert out.raw == trapdoor_proof_bytes(cs, trap, z, 5, 6)
    bad = list(z)
    bad[7] = C.r                                     # the modulus itself
    assert lib.mp_prove(h, raw(bad), raw([5]), raw([6]), out) == 1
    assert b"non-canonical" in lib.mp_last_error_detail()
    assert lib.mp_prove(h, raw(z), raw([(1 << 256) - 1]), raw([6]), out) == 1
    assert lib.mp_prove(h, raw(z), raw([5]), raw([C.r + 1]), out) == 1
    assert lib.mp_prove(h, raw(z), raw([C.r - 1]), raw([C.r - 1]), out) == 0   # the largest canonical values still prove
    assert out.raw == trapdoor_proof_bytes(cs, trap, z, C.r - 1, C.r - 1)
    ctx.close()


def test_prove_batch_is_chunked(native, monkeypatch):
    """mp_prove_batch reuses one bounded batch object over chunks of the request (here 5 proofs in chunks of 2)."""
    from manta_rs_b200 import groth16 as g16
    cs = wl.make_r1cs(3, 90, dist="R")
    pk, trap = oracle_keygen(cs, wl.sample_trapdoor(9))
    ctx = g16.ProvingContext.decode(pk)
    zs = [wl.make_assignment(cs, s) for s in range(5)]
    rs, ss = [11, 12, 13, 14, 15], [21, 22, 23, 24, 25]
    monkeypatch.setenv("MP_PROVE_BATCH_CHUNK", "2")
    proofs = g16.Groth16.prove_many_with_randomness(ctx, [g16.R1CS.from_workload(cs, z) for z in zs], rs, ss)
    for z, r, s, pr in zip(zs, rs, ss, proofs):
        assert pr.to_bytes() == trapdoor_proof_bytes(cs, trap, z, r, s)
    ctx.close()


@pytest.mark.parametrize("slabs", ["2", "4"])
def test_prove_batch_with_outer_msm_slabs(native, monkeypatch, slabs):
    """MP_ACC_SLABS: the point buffers of the bucket trees hold capacity / slabs proofs and the MSM stage runs slab after slab
    (the out-of-memory fallback of mp_batch_create takes the same path); 45 proofs = uneven slabs, every proof checked."""
    from manta_rs_b200 import groth16 as g16
    monkeypatch.setenv("MP_ACC_SLABS", slabs)
    cs = wl.make_r1cs(3, 150, dist="R")
    pk, trap = oracle_keygen(cs, wl.sample_trapdoor(15))
    ctx = g16.ProvingContext.decode(pk)
    count = 45
    zs = [wl.make_assignment(cs, 500 + s) for s in range(count)]
    rng = random.Random(int(slabs))
    rs = [rng.randrange(C.r) for _ in range(count)]
    ss = [rng.randrange(C.r) for _ in range(count)]
    proofs = g16.Groth16.prove_many_with_randomness(ctx, [g16.R1CS.from_workload(cs, z) for z in zs], rs, ss)
    for i in range(count):
        assert proofs[i].to_bytes() == trapdoor_proof_bytes(cs, trap, zs[i], rs[i], ss[i]), i
    ctx.close()


@pytest.mark.parametrize("pipe,div,count", [("1", "4", 45), ("1", "2", 70), ("1", "16", 130), ("0", "2", 45)])
def test_prove_batch_with_pipelined_tree_levels(native, monkeypatch, pipe, div, count):
    """Batches of more than 32 proofs run every tree level as a software pipeline over slabs of proofs (two round-scratch sets,
    the inversion kernels on side streams); MP_BA_PIPE_MIN_LOG2=0 forces the slab cut on a small circuit.  Uneven slabs, the
    drained and the un-pipelined forms, every proof checked against the closed form."""
    from manta_rs_b200 import groth16 as g16
    monkeypatch.setenv("MP_BA_PIPE", pipe)
    monkeypatch.setenv("MP_BA_SLAB_DIV", div)
    monkeypatch.setenv("MP_BA_PIPE_MIN_LOG2", "0")
    monkeypatch.setenv("MP_PROVE_BATCH_CHUNK", "256")
    cs = wl.make_r1cs(3, 170, dist="R")
    pk, trap = oracle_keygen(cs, wl.sample_trapdoor(21))
    ctx = g16.ProvingContext.decode(pk)
    zs = [wl.make_assignment(cs, 700 + s) for s in range(count)]
    rng = random.Random(count)
    rs = [rng.randrange(C.r) for _ in range(count)]
    ss = [rng.randrange(C.r) for _ in range(count)]
    proofs = g16.Groth16.prove_many_with_randomness(ctx, [g16.R1CS.from_workload(cs, z) for z in zs], rs, ss)
    for i in range(count):
        assert proofs[i].to_bytes() == trapdoor_proof_bytes(cs, trap, zs[i], rs[i], ss[i]), i
    ctx.close()


def test_prove_batch_more_than_one_device_batch(native):
    """260 proofs through mp_prove_batch: three device batches (128 + 128 + 4) with the default chunk."""
    from manta_rs_b200 import groth16 as g16
    cs = wl.make_r1cs(3, 200, dist="U")
    pk, trap = oracle_keygen(cs, wl.sample_trapdoor(10))
    ctx = g16.ProvingContext.decode(pk)
    count = 260
    zs = [wl.make_assignment(cs, 1000 + s) for s in range(count)]
    rng = random.Random(6)
    rs = [rng.randrange(C.r) for _ in range(count)]
    ss = [rng.randrange(C.r) for _ in range(count)]
    proofs = g16.Groth16.prove_many_with_randomness(ctx, [g16.R1CS.from_workload(cs, z) for z in zs], rs, ss)
    assert len(proofs) == count
    for i in (0, 1, 127, 128, 129, 255, 256, 259):
        assert proofs[i].to_bytes() == trapdoor_proof_bytes(cs, trap, zs[i], rs[i], ss[i]), i
    ctx.close()


# ---- Poseidon (witness-side Fr work) ----------------------------------------------------------------------------
def _poseidon_ref(state, rc, mds, width, rf_half, rp, r):
    k = 0
    for rnd in range(2 * rf_half + rp):
        state = [(x + c) % r for x, c in zip(state, rc[k:k + width])]
        k += width
        full = not (rf_half <= rnd < rf_half + rp)
        state = [pow(x, 5, r) if (full or j == 0) else x for j, x in enumerate(state)]
        state = [sum(mds[width * i + j] * state[j] for j in range(width)) % r for i in range(width)]
    return state


def test_poseidon_reference_kat_and_batch(native):
    """The reference's own known-answer vector (permutation_hardcoded_test/width3, hash.rs:248-258) through the CUDA kernel,
    then a seeded batch (widths 3 and 5, the arities manta-pay uses) against plain Python modular arithmetic."""
    gold = json.load(open(os.path.join(GOLD, "poseidon_bls381_width3.json")))
    rc = [int(x, 16) for x in gold["round_constants"]]
    mds = [int(x, 16) for x in gold["mds"]]
    lib = native.lib()
    st = ctypes.create_string_buffer(native.pack_scalars(gold["input"]), 3 * 32)
    _chk(native, lib.mp_poseidon_permute(0, 3, gold["full_rounds"], gold["partial_rounds"], native.pack_scalars(rc), native.pack_scalars(mds),
                                         st, 1, None))
    assert [str(x) for x in native.unpack_scalars(st.raw)] == gold["expected"]
    rng = random.Random(8)
    for width, rf, rp, count in ((3, 8, 55, 1000), (5, 8, 56, 300), (2, 8, 3, 5)):
        rcs = [rng.randrange(C.r) for _ in range((rf + rp) * width)] if width != 3 else rc
        m = [rng.randrange(C.r) for _ in range(width * width)] if width != 3 else mds
        states = [[rng.randrange(C.r) for _ in range(width)] for _ in range(count)]
        states[0] = [0] * width
        states[1] = [C.r - 1] * width
        buf = ctypes.create_string_buffer(native.pack_scalars([x for s in states for x in s]), count * width * 32)
        ms = ctypes.c_float()
        _chk(native, lib.mp_poseidon_permute(0, width, rf, rp, native.pack_scalars(rcs), native.pack_scalars(m), buf, count, ctypes.byref(ms)))
        got = native.unpack_scalars(buf.raw)
        for i in (0, 1, 2, count // 2, count - 1):
            assert got[i * width:(i + 1) * width] == _poseidon_ref(states[i], rcs, m, width, rf // 2, rp, C.r), (width, i)
    # host mirror of the reference's Hasher: hash(inputs) = permute(domain_tag | inputs)[0]
    from manta_rs_b200 import poseidon
    h = poseidon.Hasher(poseidon.Permutation(3, 8, 55, rc, mds), domain_tag=3)
    assert str(h.hash([1, 2])) == gold["expected"][0]
    ins = [[rng.randrange(C.r), rng.randrange(C.r)] for _ in range(50)]
    assert h.hash_many(ins) == [_poseidon_ref([3] + x, rc, mds, 3, 4, 55, C.r)[0] for x in ins]


# ---- NTT ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("log_n", [0, 1, 2, 7, 10, 11, 13, 14, 16])
def test_ntt_vs_oracle(native, log_n):
    rng = random.Random(log_n)
    n = 1 << log_n
    x = [rng.randrange(C.r) for _ in range(n)]
    x[0] = 0
    for inverse in (0, 1):
        for coset in (0, 1):
            buf = ctypes.create_string_buffer(native.pack_scalars(x), n * 32)
            _chk(native, native.lib().mp_ntt(0, buf, log_n, inverse, coset, None))
            assert native.unpack_scalars(buf.raw) == cref.ntt(x, log_n, inverse, coset), (log_n, inverse, coset)


def test_ntt_roundtrip_large(native):
    """Size-independent property at a size the oracle is not run on: ifft(fft(x)) = x, coset too."""
    log_n = 18
    n = 1 << log_n
    rng = random.Random(18)
    x = [rng.randrange(C.r) for _ in range(n)]
    packed = native.pack_scalars(x)
    for coset in (0, 1):
        buf = ctypes.create_string_buffer(packed, n * 32)
        _chk(native, native.lib().mp_ntt(0, buf, log_n, 0, coset, None))
        assert buf.raw != packed
        _chk(native, native.lib().mp_ntt(0, buf, log_n, 1, coset, None))
        assert buf.raw == packed


def test_ntt_above_2_20(native):
    """Three-pass transforms (2^21 .. 2^26 points, the size axis of kzg.rs:43-44): 2^21 against the CPU oracle, 2^22 by the
    round trip."""
    rng = random.Random(21)
    log_n = 21
    n = 1 << log_n
    raw = bytearray(rng.randbytes(n * 32))
    raw[31::32] = bytes(b & 0x3F for b in raw[31::32])          # < 2^254 < r
    x = native.unpack_scalars(bytes(raw))
    for inverse, coset in ((0, 0), (1, 1)):
        buf = ctypes.create_string_buffer(bytes(raw), n * 32)
        _chk(native, native.lib().mp_ntt(0, buf, log_n, inverse, coset, None))
        assert native.unpack_scalars(buf.raw) == cref.ntt(x, log_n, inverse, coset), (inverse, coset)
    log_n = 22
    n = 1 << log_n
    raw = bytearray(rng.randbytes(n * 32))
    raw[31::32] = bytes(b & 0x3F for b in raw[31::32])
    for coset in (0, 1):
        buf = ctypes.create_string_buffer(bytes(raw), n * 32)
        _chk(native, native.lib().mp_ntt(0, buf, log_n, 0, coset, None))
        assert buf.raw != bytes(raw)
        _chk(native, native.lib().mp_ntt(0, buf, log_n, 1, coset, None))
        assert buf.raw == bytes(raw)


# ---- full proofs -------------------------------------------------------------------------------------------------
def _prove_and_check(native, cs, pk, trap, seeds, rs, ss, oracle_full=True):
    from manta_rs_b200 import groth16 as g16
    ctx = g16.ProvingContext.decode(pk)
    zs = [wl.make_assignment(cs, s) for s in seeds]
    proofs = g16.Groth16.prove_many_with_randomness(ctx, [g16.R1CS.from_workload(cs, z) for z in zs], rs, ss)
    op = cref.OracleProver(pk, cs.p, cs.w, cs.a, cs.b, cs.c) if oracle_full else None
    for z, r, s, pr in zip(zs, rs, ss, proofs):
        assert pr.to_bytes() == trapdoor_proof_bytes(cs, trap, z, r, s)
        if op:
            assert pr.to_bytes() == op.prove(z, r, s, threads=8)
    single = g16.Groth16.prove_with_randomness(ctx, g16.R1CS.from_workload(cs, zs[0]), rs[0], ss[0])
    assert single == proofs[0]
    ctx.close()
    return proofs


@pytest.mark.parametrize("p,w,dist", [(2, 1, "U"), (2, 5, "U"), (3, 60, "R"), (5, 1200, "R"), (7, 3000, "U")])
def test_prove_small_shapes(native, p, w, dist):
    cs = wl.make_r1cs(p, w, dist=dist)
    pk, trap = oracle_keygen(cs, wl.sample_trapdoor(p + w))
    rng = random.Random(w)
    rs = [rng.randrange(C.r), 0, C.r - 1]
    ss = [rng.randrange(C.r), rng.randrange(C.r), 0]
    _prove_and_check(native, cs, pk, trap, [1, 2, 3], rs, ss)


def test_golden_proofs(native):
    from manta_rs_b200 import groth16 as g16
    gold = json.load(open(os.path.join(GOLD, "groth16_proofs.json")))
    for case in gold["cases"]:
        cs = wl.make_r1cs(case["p"], case["w"], seed=case["seed"], dist=case["dist"])
        z = wl.make_assignment(cs, case["seed"])
        pk, _ = oracle_keygen(cs, wl.sample_trapdoor(case["seed"]))
        ctx = g16.ProvingContext.decode(pk)
        pr = g16.Groth16.prove_with_randomness(ctx, g16.R1CS.from_workload(cs, z), int(case["r"]), int(case["s"]))
        assert pr.to_bytes().hex() == case["proof"]
        ctx.close()


def test_prove_with_seeded_chacha_rng_matches_explicit_randomness(native):
    """`Groth16::prove(context, compiler, rng)` draws r then s from the caller's rng (create_random_proof)."""
    from manta_rs_b200 import groth16 as g16, rng as mrng
    cs = wl.make_r1cs(3, 30)
    pk, trap = oracle_keygen(cs, wl.sample_trapdoor(4))
    z = wl.make_assignment(cs, 4)
    ctx = g16.ProvingContext.decode(pk)
    seed = bytes(range(32))
    pr = g16.Groth16.prove(ctx, g16.R1CS.from_workload(cs, z), mrng.ChaCha20Rng(seed))
    r0 = mrng.ChaCha20Rng(seed)
    r = mrng.field_rand(r0, C.r)
    s = mrng.field_rand(r0, C.r)
    assert pr.to_bytes() == trapdoor_proof_bytes(cs, trap, z, r, s)
    ctx.close()


def test_unsatisfied_assignment_matches_oracle(native):
    from manta_rs_b200 import groth16 as g16
    cs = wl.make_r1cs(2, 13)
    pk, _ = oracle_keygen(cs, wl.sample_trapdoor(2))
    z = wl.make_assignment(cs, 1)
    z[-1] = (z[-1] + 1) % C.r
    ctx = g16.ProvingContext.decode(pk)
    pr = g16.Groth16.prove_with_randomness(ctx, g16.R1CS.from_workload(cs, z), 5, 6)
    assert pr.to_bytes() == cref.OracleProver(pk, cs.p, cs.w, cs.a, cs.b, cs.c).prove(z, 5, 6)
    ctx.close()


def test_key_with_points_outside_the_subgroup_and_plain_ladder(native, monkeypatch):
    """The key is loaded like the reference's `deserialize_unchecked`: a_query / b_g1_query points that lie on the curve but
    outside the prime-order subgroup must still give ark's double-and-add result (the finishing kernel then drops its GLV
    ladder, which is only valid on the subgroup).  Also the plain ladder forced on a well-formed key."""
    from manta_rs_b200 import groth16 as g16
    from oracle.pyref.curves import Group
    G1 = Group(C, 1)
    cs = wl.make_r1cs(3, 50, dist="R")
    pk, trap = oracle_keygen(cs, wl.sample_trapdoor(31))
    rng = random.Random(31)
    outside = []
    while len(outside) < 2:
        x = rng.randrange(C.q)
        try:
            P = G1.decompress(x.to_bytes(48, "little"))
        except AssertionError:
            continue
        if G1.mul(P, C.r) is not None:
            outside.append(G1.serialize_uncompressed(P))
    pos_a = 96 + 192 * 3 + 8 + 96 * cs.p + 96 * 2 + 8
    pos_b1 = pos_a + 96 * cs.n + 8
    bad = bytearray(pk)
    bad[pos_a + 96 * 2:pos_a + 96 * 3] = outside[0]
    bad[pos_b1 + 96 * 3:pos_b1 + 96 * 4] = outside[1]
    bad = bytes(bad)
    zs = [wl.make_assignment(cs, s) for s in (1, 2)]
    rs, ss = [rng.randrange(C.r) for _ in zs], [rng.randrange(C.r) for _ in zs]
    ctx = g16.ProvingContext.decode(bad)
    op = cref.OracleProver(bad, cs.p, cs.w, cs.a, cs.b, cs.c)
    for z, r, s in zip(zs, rs, ss):
        pr = g16.Groth16.prove_with_randomness(ctx, g16.R1CS.from_workload(cs, z), r, s)
        assert pr.to_bytes() == op.prove(z, r, s, threads=4)
    ctx.close()
    monkeypatch.setenv("MP_NO_GLV", "1")
    ctx = g16.ProvingContext.decode(pk)
    for z, r, s in zip(zs, rs, ss):
        pr = g16.Groth16.prove_with_randomness(ctx, g16.R1CS.from_workload(cs, z), r, s)
        assert pr.to_bytes() == trapdoor_proof_bytes(cs, trap, z, r, s)
    ctx.close()


def test_finish_ladder_special_scalars(native):
    """r, s that stress the warp-cooperative ladders of the finishing kernel: 0, 1, 2, lambda (k1 = 0), lambda + 1, r - 1, powers
    of two, and a scalar whose GLV halves make the accumulator meet a table entry (equal operands -> complete-formula fallback)."""
    from manta_rs_b200 import groth16 as g16
    lam = 0xAC45A4010001A40200000000FFFFFFFF
    cs = wl.make_r1cs(2, 20)
    pk, trap = oracle_keygen(cs, wl.sample_trapdoor(32))
    ctx = g16.ProvingContext.decode(pk)
    z = wl.make_assignment(cs, 3)
    zsq_half = (lam + 1) // 2                      # k1 prefix = z^2 / 2 followed by bits (1, 1) with k2 = 1: acc = (1 + lambda) P meets P + phi P
    tricky = (2 * zsq_half + 1) + 1 * lam
    vals = [0, 1, 2, 3, lam, lam + 1, lam - 1, 2 * lam, C.r - 1, C.r - 2, 1 << 127, 1 << 128, (1 << 254), tricky % C.r, (lam * lam) % C.r]
    comp = g16.R1CS.from_workload(cs, z)
    rs = vals
    ss = vals[1:] + vals[:1]
    proofs = g16.Groth16.prove_many_with_randomness(ctx, [comp] * len(vals), rs, ss)
    for r, s, pr in zip(rs, ss, proofs):
        assert pr.to_bytes() == trapdoor_proof_bytes(cs, trap, z, r, s), (r, s)
    ctx.close()


def test_fwd_tma_form_matches(native):
    """The TMA-staged forward pass of the batched-affine rounds (MP_FWD_TMA=1, kept as a measured experiment) gives the same
    bytes as the default LDG form: run in a fresh process because the choice is read once per process."""
    import subprocess
    import sys
    code = (
        "import sys, ctypes, random; sys.path.insert(0, 'tests')\n"
        "from helpers import cref, BLS12_381 as C\n"
        "from manta_rs_b200 import _native as nat\n"
        "rng = random.Random(3)\n"
        "for group, n in ((1, 3000), (2, 500)):\n"
        "    pb = 96 * group\n"
        "    bases = cref.fixed_base(group, [rng.randrange(1, C.r) for _ in range(n)])\n"
        "    bases = bases[:pb * 5] + bases[pb * 4:pb * 5] * 3 + bases[pb * 8:]\n"
        "    sc = [rng.randrange(C.r) for _ in range(n)]\n"
        "    sc[5] = sc[6] = sc[7] = sc[4]\n"
        "    out = ctypes.create_string_buffer(pb)\n"
        "    fn = nat.lib().mp_msm_g1 if group == 1 else nat.lib().mp_msm_g2\n"
        "    nat.check(fn(0, bases, nat.pack_scalars(sc), n, out, None))\n"
        "    assert out.raw == cref.msm(group, bases, sc, threads=4), group\n"
        "print('ok')\n")
    env = dict(os.environ, MP_FWD_TMA="1")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-1500:]


def test_prove_from_abc(native):
    """mp_prove_from_abc: the host evaluates A z, B z, C z itself (ark `witness_map`'s first loop); same proof bytes."""
    from manta_rs_b200 import groth16 as g16
    cs = wl.make_r1cs(4, 300, dist="R")
    pk, trap = oracle_keygen(cs, wl.sample_trapdoor(14))
    z = wl.make_assignment(cs, 2)
    r_mod, m = cs.modulus, cs.m
    ev = lambda rows: [sum(c * z[j] for c, j in row) % r_mod for row in rows] + [0] * (m - cs.K)
    a, b, c = ev(cs.a), ev(cs.b), ev(cs.c)
    for j in range(cs.p):
        a[cs.K + j] = z[j]
    ctx = g16.ProvingContext.decode(pk)
    h = ctx.native(g16.R1CS.from_workload(cs, z).matrices)
    out = ctypes.create_string_buffer(192)
    P = native.pack_scalars
    _chk(native, native.lib().mp_prove_from_abc(h, P(z), P(a), P(b), P(c), P([77]), P([88]), out))
    assert out.raw == trapdoor_proof_bytes(cs, trap, z, 77, 88)
    _chk(native, native.lib().mp_prove(h, P(z), P([77]), P([88]), out))     # the matrices path still works on the same cached batch
    assert out.raw == trapdoor_proof_bytes(cs, trap, z, 77, 88)
    ctx.close()


def test_mpc_style_key_with_h_len_m(native):
    """Production keys come from the MPC and carry m (not m - 1) h_query points (mpc.rs:371-377)."""
    cs = wl.make_r1cs(3, 29)
    pk, trap = oracle_keygen(cs, wl.sample_trapdoor(6), h_len=cs.m)
    _prove_and_check(native, cs, pk, trap, [1], [77], [88])


def test_witness_map_vs_oracle(native):
    from manta_rs_b200 import groth16 as g16
    cs = wl.make_r1cs(4, 700, dist="R")
    pk, _ = oracle_keygen(cs, wl.sample_trapdoor(8))
    z = wl.make_assignment(cs, 8)
    ctx = g16.ProvingContext.decode(pk)
    h = ctx.native(g16.R1CS.from_workload(cs, z).matrices)
    out = ctypes.create_string_buffer(cs.m * 32)
    _chk(native, native.lib().mp_witness_map(h, native.pack_scalars(z), out))
    assert native.unpack_scalars(out.raw) == cref.OracleProver(pk, cs.p, cs.w, cs.a, cs.b, cs.c).witness_map(z)
    ctx.close()


@pytest.mark.parametrize("shape,dist", [("to_private", "U"), ("private_transfer", "U"), ("to_public", "R")])
def test_prove_reference_shapes_full_size(native, shape, dist):
    """BASELINE.json configs at full size: GPU keygen (checked against the oracle on a sample), batch of 3 proofs,
    every proof against the trapdoor closed form and the first against the full CPU oracle."""
    from manta_rs_b200 import groth16 as g16, keygen
    cs = wl.make_shape(shape, dist=dist)
    from oracle import trapdoor
    pk = keygen.generate(cs, wl.sample_trapdoor(21))
    _, _, trap = trapdoor.key_scalars(cs, wl.sample_trapdoor(21))   # the oracle's own QAP evaluation
    # spot-check the GPU-generated key against the oracle's fixed-base results
    r = cs.modulus
    pos_a = 96 + 192 * 3 + 8 + 96 * cs.p + 96 * 2 + 8
    idx = [0, 1, cs.n // 2, cs.n - 1]
    assert b"".join(pk[pos_a + 96 * i:pos_a + 96 * (i + 1)] for i in idx) == cref.fixed_base(1, [trap["u"][i] for i in idx])
    ctx = g16.ProvingContext.decode(pk)
    zs = [wl.make_assignment(cs, s) for s in (0, 1, 2)]
    rng = random.Random(3)
    rs = [rng.randrange(r) for _ in zs]
    ss = [rng.randrange(r) for _ in zs]
    proofs = g16.Groth16.prove_many_with_randomness(ctx, [g16.R1CS.from_workload(cs, z) for z in zs], rs, ss)
    for z, rr, s, pr in zip(zs, rs, ss, proofs):
        assert pr.to_bytes() == trapdoor_proof_bytes(cs, trap, z, rr, s)
    op = cref.OracleProver(pk, cs.p, cs.w, cs.a, cs.b, cs.c)
    assert proofs[0].to_bytes() == op.prove(zs[0], rs[0], ss[0], threads=cref.lib().oracle_max_threads())
    ctx.close()


def test_msm_closed_form_large(native):
    """2^20-point G1 MSM (BASELINE config 3 regime): bases k_i G from the fixed-base kernel, result must equal
    (sum k_i s_i) G computed by the oracle."""
    n = 1 << 20
    rng = random.Random(20)
    ks = [rng.randrange(1, C.r) for _ in range(n)]
    sc = [rng.randrange(C.r) for _ in range(n)]
    bases = ctypes.create_string_buffer(n * 96)
    _chk(native, native.lib().mp_fixed_base_g1(0, native.pack_scalars(ks), n, bases))
    sample = [0, 12345, n - 1]
    assert b"".join(bases.raw[96 * i:96 * (i + 1)] for i in sample) == cref.fixed_base(1, [ks[i] for i in sample])
    out = ctypes.create_string_buffer(96)
    _chk(native, native.lib().mp_msm_g1(0, bases, native.pack_scalars(sc), n, out, None))
    assert out.raw == cref.fixed_base(1, [sum(k * s for k, s in zip(ks, sc)) % C.r])


@pytest.mark.parametrize("env", [{}, {"MP_LADDERS_AS_MSM": "1"}, {"MP_LADDERS_AS_MSM": "1", "MP_BA_TRIM_LEVELS": "2"},
                                 {"MP_BA_TRIM_LEVELS": "1"}, {"MP_BA_TRIM_LEVELS": "4"}, {"MP_SMALL_PATH_OLD": "1"},
                                 {"MP_LADDERS_AS_MSM": "1", "MP_NO_GLV": "1"}])
def test_latency_path_of_one_and_two_proofs(native, monkeypatch, env):
    """One or two proofs (three chains on three streams).  The measured alternatives behind the environment switches give the
    same bytes: s g_a and r g1_b as two more MSM jobs over the A / B1 tables (MP_LADDERS_AS_MSM: scaled scalar vectors, lists
    sorted as 2 x count vectors) instead of ladders, bucket trees that stop MP_BA_TRIM_LEVELS levels early (the reduction adds
    the leftover points of a bucket), the first form of the enqueue order.  Capacity-2 batch with 2 and with 1 proof,
    serialised mode, special r / s."""
    from manta_rs_b200 import groth16 as g16
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    lib = native.lib()
    cs = wl.make_r1cs(3, 900, dist="R")
    pk, trap = oracle_keygen(cs, wl.sample_trapdoor(31))
    ctx = g16.ProvingContext.decode(pk)
    zs = [wl.make_assignment(cs, 40 + s) for s in range(3)]   # dist R: the boolean witnesses share one bucket (a deep tree, trimmed early)
    h = ctx.native(g16.R1CS.from_workload(cs, zs[0]).matrices)
    bt = ctypes.c_void_p()
    _chk(native, lib.mp_batch_create(h, 2, ctypes.byref(bt)))
    rng = random.Random(77)
    cases = [(rng.randrange(C.r), rng.randrange(C.r)), (0, 5), (5, 0), (C.r - 1, C.r - 1), (1, 1)]
    for overlap in (1, 0):
        _chk(native, lib.mp_batch_set_overlap(bt, overlap))
        for count in (2, 1):
            for ci, (r0, s0) in enumerate(cases):
                idx = [(ci + j) % 3 for j in range(count)]
                rs_ = [(r0 + j) % C.r for j in range(count)]
                ss_ = [(s0 + 2 * j) % C.r for j in range(count)]
                zb = native.pack_scalars([v for i in idx for v in zs[i]])
                _chk(native, lib.mp_batch_upload(bt, count, zb, native.pack_scalars(rs_), native.pack_scalars(ss_)))
                _chk(native, lib.mp_batch_run(bt, None))
                out = ctypes.create_string_buffer(count * 192)
                _chk(native, lib.mp_batch_download(bt, out))
                for j in range(count):
                    assert out.raw[j * 192:(j + 1) * 192] == trapdoor_proof_bytes(cs, trap, zs[idx[j]], rs_[j], ss_[j]), (overlap, count, ci, j)
    lib.mp_batch_destroy(bt)
    ctx.close()


def test_batch_api_sync_async_and_serialised(native):
    """The staged C ABI (upload / run / download), its asynchronous form with two batches in flight
    (submit / wait) and the serialised single-stream mode all give the oracle's bytes."""
    from manta_rs_b200 import groth16 as g16
    lib = native.lib()
    cs = wl.make_r1cs(4, 500, dist="R")
    pk, trap = oracle_keygen(cs, wl.sample_trapdoor(12))
    ctx = g16.ProvingContext.decode(pk)
    zs = [wl.make_assignment(cs, s) for s in range(6)]
    rng = random.Random(9)
    rs = [rng.randrange(C.r) for _ in zs]
    ss = [rng.randrange(C.r) for _ in zs]
    expect = [trapdoor_proof_bytes(cs, trap, z, r, s) for z, r, s in zip(zs, rs, ss)]
    h = ctx.native(g16.R1CS.from_workload(cs, zs[0]).matrices)
    batches = [ctypes.c_void_p(), ctypes.c_void_p()]
    for b in batches:
        _chk(native, lib.mp_batch_create(h, 3, ctypes.byref(b)))
    pack = lambda lo, hi: (native.pack_scalars([v for z in zs[lo:hi] for v in z]), native.pack_scalars(rs[lo:hi]), native.pack_scalars(ss[lo:hi]))
    # synchronous staged form, overlapped and serialised
    for overlap in (1, 0):
        _chk(native, lib.mp_batch_set_overlap(batches[0], overlap))
        zb, rb, sb = pack(0, 3)
        _chk(native, lib.mp_batch_upload(batches[0], 3, zb, rb, sb))
        ms = ctypes.c_float()
        _chk(native, lib.mp_batch_run(batches[0], ctypes.byref(ms)))
        out = ctypes.create_string_buffer(3 * 192)
        _chk(native, lib.mp_batch_download(batches[0], out))
        assert [out.raw[i * 192:(i + 1) * 192] for i in range(3)] == expect[0:3]
        assert ms.value > 0 and lib.mp_batch_kernel_launches(batches[0]) > 10
    _chk(native, lib.mp_batch_set_overlap(batches[0], 1))
    # two batches in flight
    bufs = [pack(0, 3), pack(3, 6)]
    outs = [ctypes.create_string_buffer(3 * 192), ctypes.create_string_buffer(3 * 192)]
    for i in range(2):
        _chk(native, lib.mp_batch_submit(batches[i], 3, bufs[i][0], bufs[i][1], bufs[i][2], outs[i]))
    # a batch that is in flight refuses new work instead of corrupting it
    assert lib.mp_batch_run_async(batches[0]) == 1
    for i in range(2):
        _chk(native, lib.mp_batch_wait(batches[i], None))
    assert [outs[0].raw[i * 192:(i + 1) * 192] for i in range(3)] == expect[0:3]
    assert [outs[1].raw[i * 192:(i + 1) * 192] for i in range(3)] == expect[3:6]
    # partial batch and empty batch
    zb, rb, sb = pack(4, 5)
    _chk(native, lib.mp_batch_upload(batches[1], 1, zb, rb, sb))
    _chk(native, lib.mp_batch_run(batches[1], None))
    one = ctypes.create_string_buffer(192)
    _chk(native, lib.mp_batch_download(batches[1], one))
    assert one.raw == expect[4]
    assert lib.mp_batch_upload(batches[1], 4, zb, rb, sb) == 1     # over capacity
    for b in batches:
        lib.mp_batch_destroy(b)
    ctx.close()
