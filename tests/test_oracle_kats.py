"""Pins the CPU oracle against the reference's own known-answer fixtures — no GPU needed."""
import json
import os

from helpers import cref, BLS12_381
from oracle.pyref.fields import BN254, CURVES
from oracle.pyref.curves import Group
from oracle.pyref import pairing

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _poseidon(perm_mul, perm_add, gold):
    """x^5 Poseidon, width 3, 8 full + 55 partial rounds (reference KAT `hash.rs:248-258`)."""
    rc = [int(x, 16) for x in gold["round_constants"]]
    mds = [int(x, 16) for x in gold["mds"]]
    state = list(gold["input"])
    rf, rp = gold["full_rounds"] // 2, gold["partial_rounds"]
    k = 0

    def sbox(x):
        x2 = perm_mul([x], [x])[0]
        x4 = perm_mul([x2], [x2])[0]
        return perm_mul([x4], [x])[0]

    def mix(st):
        out = []
        for i in range(3):
            prods = perm_mul(mds[3 * i:3 * i + 3], st)
            acc = prods[0]
            for v in prods[1:]:
                acc = perm_add([acc], [v])[0]
            out.append(acc)
        return out

    for rnd in range(2 * rf + rp):
        state = perm_add(state, rc[k:k + 3])
        k += 3
        if rf <= rnd < rf + rp:
            state[0] = sbox(state[0])
        else:
            state = [sbox(x) for x in state]
        state = mix(state)
    return state


def test_poseidon_kat_python_oracle():
    gold = json.load(open(os.path.join(GOLD, "poseidon_bls381_width3.json")))
    r = BLS12_381.r
    out = _poseidon(lambda a, b: [x * y % r for x, y in zip(a, b)], lambda a, b: [(x + y) % r for x, y in zip(a, b)], gold)
    assert [str(x) for x in out] == gold["expected"]


def test_poseidon_kat_cpp_oracle():
    gold = json.load(open(os.path.join(GOLD, "poseidon_bls381_width3.json")))
    out = _poseidon(lambda a, b: cref.field_op(1, 2, a, b), lambda a, b: cref.field_op(1, 0, a, b), gold)
    assert [str(x) for x in out] == gold["expected"]


def test_bn254_verifying_key_pairing_kat():
    """Decompress alpha_g1 / beta_g2 from the reference's checked-in verifying keys and reproduce the stored
    e(alpha, beta) bytes: pins the compressed-point flag conventions and the Fq2 / Fq12 serialization order."""
    gold = json.load(open(os.path.join(GOLD, "bn254_vk_kat.json")))
    G1, G2 = Group(BN254, 1), Group(BN254, 2)
    expected_sizes = {"to-private": (35994, 13), "private-transfer": (36442, 27), "to-public": (36186, 19)}
    for name, kat in gold.items():
        assert (kat["file_size"], kat["gamma_abc_len"]) == expected_sizes[name]
        alpha = G1.decompress(bytes.fromhex(kat["alpha_g1"]))
        beta = G2.decompress(bytes.fromhex(kat["beta_g2"]))
        assert G1.on_curve(alpha) and G2.on_curve(beta)
        assert G1.compress(alpha).hex() == kat["alpha_g1"] and G2.compress(beta).hex() == kat["beta_g2"]
        # MPC keys fix gamma to the G2 generator (manta-trusted-setup/src/groth16/mpc.rs:419)
        assert G2.decompress(bytes.fromhex(kat["gamma_g2"])) == BN254.g2
        for hx in kat["gamma_abc_g1"]:
            assert G1.on_curve(G1.decompress(bytes.fromhex(hx)))
        e = pairing.pairing_bn254_ark(BN254, alpha, beta)
        assert pairing.fq12_to_bytes(BN254, e).hex() == kat["alpha_g1_beta_g2"], name
        # the sign flag matters: flipping y of either input must NOT reproduce the stored value
        e_neg = pairing.pairing_bn254_ark(BN254, G1.neg(alpha), beta)
        assert pairing.fq12_to_bytes(BN254, e_neg).hex() != kat["alpha_g1_beta_g2"]


def test_curve_constants_self_check():
    for curve in CURVES.values():
        G1, G2 = Group(curve, 1), Group(curve, 2)
        assert G1.on_curve(curve.g1) and G2.on_curve(curve.g2)
        assert G1.mul(curve.g1, curve.r) is None and G2.mul(curve.g2, curve.r) is None
        w = curve.root_of_unity
        assert pow(w, 1 << curve.two_adicity, curve.r) == 1 and pow(w, 1 << (curve.two_adicity - 1), curve.r) == curve.r - 1
    # Appendix A: the G1 generator's y is the smaller of {y, -y} -> flag 0x00
    assert Group(BLS12_381, 1).compress(BLS12_381.g1)[-1] & 0x80 == 0
    assert BLS12_381.root_of_unity == 10238227357739495823651030575849232062558860180284477541189508159991286009131
