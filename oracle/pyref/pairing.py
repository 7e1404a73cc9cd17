"""TEST INFRASTRUCTURE ONLY — CPU oracle: pairings, used only to pin conventions against the reference's fixtures.

* BN254 optimal ate, reproducing the value ark-ec 0.3 `Bn<P>::pairing` stores in the reference's prepared verifying
  keys (`manta-parameters/data/pay/verifying/*.dat`, field `alpha_g1_beta_g2`, `groth16.rs:338-356`): ark's final
  exponentiation computes f^((q^12-1)/r * 2x(6x^2+3x+1)) (the Fuentes-Castaneda hard part), so that multiple is
  applied here.  Matching those 384 bytes pins the y-sign flag of compressed G1 AND G2 points (a wrong sign on
  either input inverts the pairing value) and the Fq2/Fq6/Fq12 serialization order.
* BLS12-381 ate pairing (any fixed power of the reduced pairing) for verify-equation self-checks of the oracle.

Line functions are evaluated on the twist with Fq2 slopes and embedded sparsely into Fq12 = Fq6[w]/(w^2 - v).
"""
from __future__ import annotations

from .fields import Tower, fq2_add, fq2_sub, fq2_mul, fq2_sqr, fq2_inv, fq2_neg, fq2_conj, fq2_pow, fq2_scalar


def _line(curve, tw: Tower, T, Q, P):
    """Line through twist points T, Q (tangent when equal) evaluated at P in G1; returns (sparse Fq12, T + Q)."""
    q = curve.q
    xt, yt = T
    xq, yq = Q
    if T == Q:
        lam = fq2_mul(q, fq2_scalar(q, fq2_sqr(q, xt), 3), fq2_inv(q, fq2_scalar(q, yt, 2)))
    else:
        lam = fq2_mul(q, fq2_sub(q, yq, yt), fq2_inv(q, fq2_sub(q, xq, xt)))
    x3 = fq2_sub(q, fq2_sub(q, fq2_sqr(q, lam), xt), xq)
    y3 = fq2_sub(q, fq2_mul(q, lam, fq2_sub(q, xt, x3)), yt)
    xp, yp = P
    a = fq2_neg(q, fq2_scalar(q, lam, xp))                    # -lambda * xP
    b = fq2_sub(q, fq2_mul(q, lam, xt), yt)                   # lambda * xT - yT
    zero = (0, 0)
    if curve.twist_type == 'D':
        # l = yP - (lambda xP) w + (lambda xT - yT) w^3,  w^3 = v w
        ell = (((yp % q, 0), zero, zero), (a, b, zero))
    else:
        # M twist: l * w^3 = (lambda xT - yT) - (lambda xP) w^2 + yP w^3   (w^3 lies in a proper subfield)
        ell = ((b, a, zero), (zero, (yp % q, 0), zero))
    return ell, (x3, y3)


def _frobenius_twist(curve, Q, power):
    """pi^power on D-twist coordinates: (conj^power(x) * xi^((q^power-1)/3), conj^power(y) * xi^((q^power-1)/2))."""
    q = curve.q
    x, y = Q
    for _ in range(power):
        x = fq2_mul(q, fq2_conj(q, x), fq2_pow(q, curve.xi, (q - 1) // 3))
        y = fq2_mul(q, fq2_conj(q, y), fq2_pow(q, curve.xi, (q - 1) // 2))
    return (x, y)


def miller_loop_bn(curve, P, Q):
    tw = Tower(curve.q, curve.xi)
    if P is None or Q is None:
        return tw.one12
    n = 6 * curve.x + 2
    f = tw.one12
    T = Q
    for bit in bin(n)[3:]:
        ell, T2 = _line(curve, tw, T, T, P)
        f = tw.mul12(tw.sqr12(f), ell)
        T = T2
        if bit == "1":
            ell, T2 = _line(curve, tw, T, Q, P)
            f = tw.mul12(f, ell)
            T = T2
    Q1 = _frobenius_twist(curve, Q, 1)
    Q2 = _frobenius_twist(curve, Q, 2)
    Q2 = (Q2[0], fq2_neg(curve.q, Q2[1]))
    ell, T2 = _line(curve, tw, T, Q1, P)
    f = tw.mul12(f, ell)
    T = T2
    ell, _ = _line(curve, tw, T, Q2, P)
    f = tw.mul12(f, ell)
    return f


def pairing_bn254_ark(curve, P, Q):
    """e(P, Q) exactly as ark-ec 0.3 computes it for BN curves (see module docstring)."""
    tw = Tower(curve.q, curve.xi)
    f = miller_loop_bn(curve, P, Q)
    x = curve.x
    k = 2 * x * (6 * x * x + 3 * x + 1)
    e = (curve.q ** 12 - 1) // curve.r * k
    return tw.pow12(f, e)


def miller_loop_bls(curve, P, Q):
    tw = Tower(curve.q, curve.xi)
    if P is None or Q is None:
        return tw.one12
    f = tw.one12
    T = Q
    for bit in bin(curve.x)[3:]:
        ell, T2 = _line(curve, tw, T, T, P)
        f = tw.mul12(tw.sqr12(f), ell)
        T = T2
        if bit == "1":
            ell, T2 = _line(curve, tw, T, Q, P)
            f = tw.mul12(f, ell)
            T = T2
    return f


def pairing_bls(curve, P, Q):
    """A bilinear non-degenerate pairing on BLS12-381 (ate, x taken positive; a fixed power of ark's value)."""
    tw = Tower(curve.q, curve.xi)
    return tw.pow12(miller_loop_bls(curve, P, Q), (curve.q ** 12 - 1) // curve.r)


def fq12_to_bytes(curve, f) -> bytes:
    """ark-serialize of Fq12: c0 (Fq6: c0, c1, c2 as Fq2: c0, c1) then c1; little-endian field elements."""
    n = curve.fq_ser_bytes
    out = bytearray()
    for half in f:
        for c in half:
            out += c[0].to_bytes(n, "little") + c[1].to_bytes(n, "little")
    return bytes(out)


def groth16_verify(curve, vk, public_inputs, proof):
    """e(A, B) = e(alpha, beta) * e(sum x_i gamma_abc_i, gamma) * e(C, delta)   (SURVEY.md C.7).
    public_inputs excludes the leading constant 1."""
    from .curves import Group
    G1 = Group(curve, 1)
    tw = Tower(curve.q, curve.xi)
    pair = pairing_bn254_ark if curve.name == "bn254" else pairing_bls
    acc = G1.to_jac(vk["gamma_abc_g1"][0])
    for x, pt in zip(public_inputs, vk["gamma_abc_g1"][1:]):
        acc = G1.jac_add(acc, G1.jac_mul(G1.to_jac(pt), x % curve.r))
    lhs = pair(curve, proof[0], proof[1])
    rhs = tw.mul12(tw.mul12(pair(curve, vk["alpha_g1"], vk["beta_g2"]), pair(curve, G1.to_affine(acc), vk["gamma_g2"])),
                   pair(curve, proof[2], vk["delta_g2"]))
    return lhs == rhs
