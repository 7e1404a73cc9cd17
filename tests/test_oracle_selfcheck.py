"""The two oracle implementations (Python big-int, C++) against each other, against the trapdoor closed form and
against the committed golden proofs; plus the verify equation through the oracle's pairing."""
import json
import os
import random

from helpers import cref, BLS12_381 as C, oracle_keygen, trapdoor_proof_bytes
from oracle.pyref import groth16 as og, curves, pairing
from oracle.pyref.poly import Radix2Domain
import manta_rs_b200.workload as wl

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_field_ops_cpp_vs_python():
    rng = random.Random(3)
    for f, p in ((0, C.q), (1, C.r)):
        a = [rng.randrange(p) for _ in range(64)] + [0, 1, p - 1]
        b = [rng.randrange(p) for _ in range(64)] + [p - 1, 0, p - 1]
        assert cref.field_op(f, 0, a, b) == [(x + y) % p for x, y in zip(a, b)]
        assert cref.field_op(f, 1, a, b) == [(x - y) % p for x, y in zip(a, b)]
        assert cref.field_op(f, 2, a, b) == [x * y % p for x, y in zip(a, b)]
        assert cref.field_op(f, 4, a) == [pow(x, -1, p) if x else 0 for x in a]
        assert cref.field_op(f, 5, a) == [(-x) % p for x in a]


def test_msm_cpp_vs_python_pippenger():
    rng = random.Random(4)
    for gid in (1, 2):
        G = curves.Group(C, gid)
        n = 45
        ks = [rng.randrange(C.r) for _ in range(n)]
        fb = cref.fixed_base(gid, ks)
        pb = 96 * gid
        pts = [G.deserialize_uncompressed(fb[i * pb:(i + 1) * pb]) for i in range(n)]
        assert pts[:4] == [G.mul(G.gen, k) for k in ks[:4]]
        sc = [rng.randrange(C.r) for _ in range(n)]
        sc[3], sc[4], sc[5] = 0, 1, C.r - 1
        got = G.deserialize_uncompressed(cref.msm(gid, fb, sc))
        assert got == G.to_affine(curves.msm_pippenger(G, pts, sc))
        assert got == G.to_affine(curves.msm_naive(G, pts, sc))
        assert cref.msm(gid, fb, sc, threads=4) == cref.msm(gid, fb, sc)


def test_ntt_cpp_vs_python():
    rng = random.Random(5)
    for logn in (0, 1, 2, 5, 9):
        n = 1 << logn
        d = Radix2Domain(C, n)
        x = [rng.randrange(C.r) for _ in range(n)]
        assert cref.ntt(x, logn, 0, 0) == d.fft(x)
        assert cref.ntt(x, logn, 1, 0) == d.ifft(x)
        assert cref.ntt(x, logn, 0, 1) == d.coset_fft(x)
        assert cref.ntt(x, logn, 1, 1) == d.coset_ifft(x)
        assert d.ifft(d.fft(x)) == x and d.coset_ifft(d.coset_fft(x)) == x


def test_prove_cpp_vs_python_vs_trapdoor_and_verify():
    cs = wl.make_r1cs(3, 24, dist="R")
    z = wl.make_assignment(cs, 2)
    assert wl.is_satisfied(cs, z)
    pk, trap = og.setup_trapdoor(C, cs.as_dict(), *wl.sample_trapdoor(9))
    pkb = og.pk_to_bytes(C, pk)
    # the CPU key generator used by the tests produces the same bytes as the Python restatement of ark's generator
    assert oracle_keygen(cs, wl.sample_trapdoor(9))[0] == pkb
    assert og.pk_from_bytes(C, pkb) == pk
    op = cref.OracleProver(pkb, cs.p, cs.w, cs.a, cs.b, cs.c)
    assert op.witness_map(z) == og.witness_map(C, cs.as_dict(), z)[0]
    for r, s in ((123, 456), (0, 5), (C.r - 1, C.r - 2)):
        proof = op.prove(z, r, s)
        assert proof == og.proof_to_bytes(C, og.create_proof(C, pk, cs.as_dict(), z, r, s))
        assert proof == og.proof_to_bytes(C, og.trapdoor_proof(C, cs.as_dict(), trap, z, r, s))
        assert proof == trapdoor_proof_bytes(cs, trap, z, r, s)
        assert proof == op.prove(z, r, s, threads=4)
    # Groth16 verification equation on the last proof (oracle pairing), and rejection of a wrong public input
    pr = og.proof_from_bytes(C, proof)
    assert pairing.groth16_verify(C, pk["vk"], z[1:cs.p], pr)
    bad = list(z[1:cs.p])
    bad[0] = (bad[0] + 1) % C.r
    assert not pairing.groth16_verify(C, pk["vk"], bad, pr)


def test_golden_proofs_cpp_oracle():
    gold = json.load(open(os.path.join(GOLD, "groth16_proofs.json")))
    for case in gold["cases"]:
        cs = wl.make_r1cs(case["p"], case["w"], seed=case["seed"], dist=case["dist"])
        z = wl.make_assignment(cs, case["seed"])
        pk, _ = oracle_keygen(cs, wl.sample_trapdoor(case["seed"]))
        op = cref.OracleProver(pk, cs.p, cs.w, cs.a, cs.b, cs.c)
        assert op.prove(z, int(case["r"]), int(case["s"])).hex() == case["proof"]


def test_unsatisfied_assignment_still_deterministic():
    """An assignment that violates a constraint gives h with a non-zero top coefficient; the MSM silently
    truncates it against h_query (m - 1 points) exactly like ark."""
    cs = wl.make_r1cs(2, 13, dist="U")
    z = wl.make_assignment(cs, 1)
    z[-1] = (z[-1] + 1) % C.r
    assert not wl.is_satisfied(cs, z)
    pk, _ = og.setup_trapdoor(C, cs.as_dict(), *wl.sample_trapdoor(2))
    op = cref.OracleProver(og.pk_to_bytes(C, pk), cs.p, cs.w, cs.a, cs.b, cs.c)
    assert op.prove(z, 5, 6) == og.proof_to_bytes(C, og.create_proof(C, pk, cs.as_dict(), z, 5, 6))


def test_glv_constants_of_the_finishing_kernel():
    """The endomorphism constants generated into the CUDA constants header (tools/gen_constants.py): lambda = z^2 - 1 is a
    primitive cube root of unity mod r, and phi(x, y) = (beta x, y) equals lambda (x, y) on G1 for the beta that is emitted."""
    from oracle.pyref.fields import BLS12_381 as C
    from oracle.pyref.curves import Group
    G = Group(C, 1)
    z = -0xd201000000010000
    lam = (z * z - 1) % C.r
    beta = pow(pow(2, (C.q - 1) // 3, C.q), 2, C.q)
    assert (lam * lam + lam + 1) % C.r == 0 and lam.bit_length() == 128
    assert beta != 1 and pow(beta, 3, C.q) == 1
    for k in (1, 5, 123456789):
        P = G.mul(G.gen, k)
        assert G.mul(P, lam) == (beta * P[0] % C.q, P[1])
    # every scalar splits as k2 * lambda + k1 with both halves below 2^128
    assert ((C.r - 1) // lam).bit_length() <= 128
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "manta-rs_b200", "csrc", "bls12_381_constants.cuh")).read()
    mont = beta * (1 << 384) % C.q
    assert "FQ_GLV_BETA[12] = {" + ", ".join("0x%08xu" % ((mont >> (32 * i)) & 0xFFFFFFFF) for i in range(12)) + "}" in hdr


def test_shared_inversion_of_g2_chords_through_norms():
    """The algebra of the G2 batched-affine rounds (msm_impl.inc BaInv<Fq2>): the denominators x2 - x1 of a set of
    Fq2 chord additions are inverted through ONE Fq inversion of the product of their norms - forward prefix products of
    a^2 + b^2, back-substitution, 1/d = conj(d) * n^-1 - and the chord formula with those inverses equals the group law."""
    import random
    from oracle.pyref.fields import BLS12_381 as C, fq2_sub, fq2_mul, fq2_sqr, fq2_conj, fq2_scalar
    from oracle.pyref.curves import Group
    G = Group(C, 2)
    q = C.q
    rng = random.Random(41)
    pts = [G.mul(G.gen, rng.randrange(1, C.r)) for _ in range(12)]
    pairs = [(pts[2 * i], pts[2 * i + 1]) for i in range(6)]
    d = [fq2_sub(q, b[0], a[0]) for a, b in pairs]
    norms = [(x[0] * x[0] + x[1] * x[1]) % q for x in d]
    prefix, run = [], 1
    for n in norms:                       # k_ba_fwd: running product of the norms
        run = run * n % q
        prefix.append(run)
    inv = pow(run, -1, q)                 # k_ba_mid: the only inversion, in Fq
    for t in range(len(pairs) - 1, -1, -1):   # k_ba_bwd
        ninv = inv * prefix[t - 1] % q if t else inv
        inv = inv * norms[t] % q
        dinv = fq2_scalar(q, fq2_conj(q, d[t]), ninv)
        assert fq2_mul(q, dinv, d[t]) == (1, 0)
        (x1, y1), (x2, y2) = pairs[t]
        lam = fq2_mul(q, fq2_sub(q, y2, y1), dinv)
        x3 = fq2_sub(q, fq2_sub(q, fq2_sqr(q, lam), x1), x2)
        y3 = fq2_sub(q, fq2_mul(q, lam, fq2_sub(q, x1, x3)), y1)
        assert (x3, y3) == G.add(pairs[t][0], pairs[t][1])
