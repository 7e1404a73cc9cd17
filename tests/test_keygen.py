"""Key generation: the oracle's restatement of the trusted-setup `initialize` against the known-trapdoor
key on CPU, and the device kernels (`mp_keygen`, `mp_mpc_initialize`, `mp_group_ntt`) against both on the GPU."""
import random

import pytest

from helpers import cref, BLS12_381 as C, oracle_keygen
import manta_rs_b200.workload as wl


def dummy_circuit():
    """`dummy_circuit` of manta-trusted-setup/src/groth16/test/mod.rs:209-217: secret a = 2, b = 3, c = a * b, public d = 6,
    c == d.  Variables: [1, d | a, b, c]; rows: a * b = c ; (c - d) * 1 = 0."""
    r = C.r
    cs = wl.R1CS(modulus=r, p=2, w=3, K=2)
    cs.a = [[(1, 2)], [(1, 4), (r - 1, 1)]]
    cs.b = [[(1, 3)], [(1, 0)]]
    cs.c = [[(1, 4)], []]
    cs.kinds = [0, 0]
    return cs, [1, 6, 2, 3, 6]


def phase1_powers(m, tau, alpha, beta):
    """The phase-1 accumulator (kzg.rs:443-470 shape) for known secrets, as ark uncompressed bytes."""
    r = C.r
    tp = [pow(tau, i, r) for i in range(2 * m)]
    tau1 = cref.fixed_base(1, tp)
    tau2 = cref.fixed_base(2, tp[:m])
    alpha1 = cref.fixed_base(1, [alpha * t % r for t in tp[:m]])
    beta1 = cref.fixed_base(1, [beta * t % r for t in tp[:m]])
    beta2 = cref.fixed_base(2, [beta])
    return tau1, tau2, alpha1, beta1, beta2


def test_oracle_initialize_matches_trapdoor_key_on_dummy_circuit():
    """mpc.rs:355-431 restated in oracle/pyref/mpc.py: on the reference's dummy circuit (and a slightly larger system) the
    initial MPC state built from phase-1 powers of known secrets equals the known-trapdoor key with gamma = delta = 1 and
    m h_query points - the identity that also pins the device kernel at sizes the Python restatement cannot reach."""
    from oracle.pyref import mpc, groth16 as og
    from oracle.pyref.curves import Group
    G1, G2 = Group(C, 1), Group(C, 2)
    for cs, tag in ((dummy_circuit()[0], "dummy"), (wl.make_r1cs(2, 5, seed=3), "w5")):
        tau, alpha, beta = 0x1234567, 0x89ABCDE, 0xF0F0F0F1
        m = cs.m
        tau1, tau2, alpha1, beta1, beta2 = phase1_powers(m, tau, alpha, beta)
        de1 = lambda buf: [G1.deserialize_uncompressed(buf[96 * i:96 * i + 96]) for i in range(len(buf) // 96)]
        de2 = lambda buf: [G2.deserialize_uncompressed(buf[192 * i:192 * i + 192]) for i in range(len(buf) // 192)]
        powers = dict(tau_powers_g1=de1(tau1), tau_powers_g2=de2(tau2), alpha_tau_powers_g1=de1(alpha1), beta_tau_powers_g1=de1(beta1),
                      beta_g2=de2(beta2)[0])
        pk = mpc.initialize(C, powers, cs.as_dict())
        want, _ = oracle_keygen(cs, (tau, alpha, beta, 1, 1), h_len=m)
        assert og.pk_to_bytes(C, pk) == want, tag


@pytest.mark.gpu
@pytest.mark.parametrize("p,w,dist,h_m", [(2, 1, "U", False), (3, 60, "R", True), (5, 700, "R", False)])
def test_device_keygen_matches_oracle_key(native, p, w, dist, h_m):
    from manta_rs_b200 import keygen
    cs = wl.make_r1cs(p, w, dist=dist)
    trap = wl.sample_trapdoor(p + w)
    want, _ = oracle_keygen(cs, trap, h_len=cs.m if h_m else None)
    assert keygen.generate(cs, trap, h_len=cs.m if h_m else None) == want


@pytest.mark.gpu
def test_device_keygen_rejects_bad_trapdoor(native):
    from manta_rs_b200 import keygen
    cs = wl.make_r1cs(2, 5)
    omega = pow(pow(7, (C.r - 1) >> 32, C.r), 1 << (32 - cs.log_m), C.r)
    for trap in ((0, 2, 3, 4, 5), (2, 3, 4, 5, C.r), (omega, 2, 3, 4, 5)):      # zero, non-canonical, tau inside the domain
        with pytest.raises(native.NativeError):
            keygen.generate(cs, trap)


@pytest.mark.gpu
@pytest.mark.parametrize("group", [1, 2])
def test_group_ntt_vs_oracle_and_roundtrip(native, group):
    from manta_rs_b200 import keygen
    from oracle.pyref import mpc
    from oracle.pyref.curves import Group
    from oracle.pyref.poly import Radix2Domain
    G = Group(C, group)
    pb = 96 * group
    rng = random.Random(group)
    for log_n in (0, 1, 3):
        n = 1 << log_n
        pts = bytearray(cref.fixed_base(group, [rng.randrange(1, C.r) for _ in range(n)]))
        if n > 2:
            pts[pb:2 * pb] = G.serialize_uncompressed(None)
        pts = bytes(pts)
        dom = Radix2Domain(C, n)
        aff = [G.deserialize_uncompressed(pts[pb * i:pb * i + pb]) for i in range(n)]
        for inverse in (False, True):
            want = b"".join(G.serialize_uncompressed(P) for P in mpc.group_fft(G, aff, dom, inverse))
            assert keygen.group_ntt(group, pts, inverse) == want, (group, log_n, inverse)
    # size-independent property at a size the Python oracle is not run on: ifft(fft(x)) == x, and linearity against the
    # Fr transform: fft of k_i * G equals (fft of k)_i * G
    log_n = 9 if group == 1 else 7
    n = 1 << log_n
    ks = [rng.randrange(C.r) for _ in range(n)]
    pts = cref.fixed_base(group, ks)
    fwd = keygen.group_ntt(group, pts, False)
    assert fwd == cref.fixed_base(group, cref.ntt(ks, log_n, 0, 0))
    assert keygen.group_ntt(group, fwd, True) == pts


@pytest.mark.gpu
def test_mpc_initialize_dummy_circuit_and_prove(native):
    """The reference's own trusted-setup fixture (test/mod.rs:209-217, 263-286): initialize on the dummy circuit 2 * 3 = 6,
    checked against the oracle's restatement of mpc.rs:355-431, then a proof with the resulting key verifies (pairing check)."""
    from manta_rs_b200 import keygen, groth16 as g16
    from oracle.pyref import mpc, groth16 as og, pairing
    from oracle.pyref.curves import Group
    G1, G2 = Group(C, 1), Group(C, 2)
    cs, z = dummy_circuit()
    tau, alpha, beta = 0x1234567, 0x89ABCDE, 0xF0F0F0F1
    tau1, tau2, alpha1, beta1, beta2 = phase1_powers(cs.m, tau, alpha, beta)
    pk = keygen.mpc_initialize(cs, tau1, tau2, alpha1, beta1, beta2)
    de1 = lambda buf: [G1.deserialize_uncompressed(buf[96 * i:96 * i + 96]) for i in range(len(buf) // 96)]
    de2 = lambda buf: [G2.deserialize_uncompressed(buf[192 * i:192 * i + 192]) for i in range(len(buf) // 192)]
    powers = dict(tau_powers_g1=de1(tau1), tau_powers_g2=de2(tau2), alpha_tau_powers_g1=de1(alpha1), beta_tau_powers_g1=de1(beta1),
                  beta_g2=de2(beta2)[0])
    assert pk == og.pk_to_bytes(C, mpc.initialize(C, powers, cs.as_dict()))
    ctx = g16.ProvingContext.decode(pk)
    proof = g16.Groth16.prove_with_randomness(ctx, g16.R1CS.from_workload(cs, z), 0xABCDEF, 0x13579B)
    ctx.close()
    key = og.pk_from_bytes(C, pk)
    proof_pts = og.proof_from_bytes(C, proof.to_bytes())
    assert pairing.groth16_verify(C, key["vk"], [6], proof_pts)
    assert not pairing.groth16_verify(C, key["vk"], [7], proof_pts)


@pytest.mark.gpu
@pytest.mark.parametrize("p,w", [(3, 200), (13, 8240)])
def test_mpc_initialize_vs_trapdoor_key(native, p, w):
    """At sizes the Python restatement cannot reach (up to the ToPrivate shape, m = 2^14): the device `initialize` on powers of
    known secrets equals the known-trapdoor key with gamma = delta = 1 (identity pinned on CPU by the test above)."""
    from manta_rs_b200 import keygen
    cs = wl.make_r1cs(p, w, dist="R")
    tau, alpha, beta = wl.sample_trapdoor(77)[:3]
    tau1, tau2, alpha1, beta1, beta2 = phase1_powers(cs.m, tau, alpha, beta)
    want, _ = oracle_keygen(cs, (tau, alpha, beta, 1, 1), h_len=cs.m)
    assert keygen.mpc_initialize(cs, tau1, tau2, alpha1, beta1, beta2) == want
