"""TEST INFRASTRUCTURE ONLY — known-trapdoor Groth16 keys and the closed form of a proof.

Follows the key structure ark-groth16 0.3 `generator.rs` produces (un-vendored; the in-repo re-derivation is
`manta-trusted-setup/src/groth16/mpc.rs:245-312,355-431`): with toxic waste (tau, alpha, beta, gamma, delta) and
u_i, v_i, w_i the QAP polynomials of variable i evaluated at tau, the proof for (z, r, s) is
    A = (alpha + sum z_i u_i + r delta) G1,   B = (beta + sum z_i v_i + s delta) G2,
    C = (sum_{i>=p} z_i (beta u_i + alpha v_i + w_i) / delta + h(tau) Z(tau) / delta + s A_s + r B_s - r s delta) G1
— Fr arithmetic on Python integers plus three fixed-base multiplications by the C++ oracle: an answer that is
independent of every MSM / NTT code path of the product.  Only tests/, __graft_entry__.smoke() and bench.py's
checking legs may import this module; the product package never does.
"""
from __future__ import annotations

from operator import mul


def _batch_inv(vals, r):
    prods, acc = [], 1
    for v in vals:
        acc = acc * v % r
        prods.append(acc)
    inv = pow(acc, -1, r)
    out = [0] * len(vals)
    for i in range(len(vals) - 1, -1, -1):
        out[i] = inv * (prods[i - 1] if i else 1) % r
        inv = inv * vals[i] % r
    return out


def lagrange_at_tau(modulus, log_m, tau):
    """L_j(tau) = Z(tau)/m * w^j / (tau - w^j) over the radix-2 domain of size 2^log_m (generator 7, 2-adicity 32)."""
    r, m = modulus, 1 << log_m
    root = pow(7, (r - 1) >> 32, r)
    omega = pow(root, 1 << (32 - log_m), r)
    zt = (pow(tau, m, r) - 1) % r
    assert zt != 0
    ws, wj = [], 1
    for _ in range(m):
        ws.append(wj)
        wj = wj * omega % r
    invs = _batch_inv([(tau - x) % r for x in ws], r)
    zm = zt * pow(m, -1, r) % r
    return [zm * x % r * y % r for x, y in zip(ws, invs)], zt


def qap_at_tau(cs, tau):
    """u_i(tau), v_i(tau), w_i(tau) for all variables and Z(tau); `cs` is a workload.R1CS (rows of (coeff, column))."""
    r = cs.modulus
    L, zt = lagrange_at_tau(r, cs.log_m, tau)
    n, p, K = cs.n, cs.p, cs.K
    u, v, w = [0] * n, [0] * n, [0] * n
    for i in range(p):
        u[i] = L[K + i]
    for j in range(K):
        lj = L[j]
        for coeff, col in cs.a[j]:
            u[col] = (u[col] + lj * coeff) % r
        for coeff, col in cs.b[j]:
            v[col] = (v[col] + lj * coeff) % r
        for coeff, col in cs.c[j]:
            w[col] = (w[col] + lj * coeff) % r
    return u, v, w, zt


def key_scalars(cs, trapdoor, h_len=None):
    """Discrete logs of every proving-key element: (g1 scalars in file order pieces, g2 scalars, trap dict)."""
    tau, alpha, beta, gamma, delta = trapdoor
    r = cs.modulus
    u, v, w, zt = qap_at_tau(cs, tau)
    n, p, m = cs.n, cs.p, cs.m
    h_len = m - 1 if h_len is None else h_len
    ginv, dinv = pow(gamma, -1, r), pow(delta, -1, r)
    abc = [(beta * u[i] + alpha * v[i] + w[i]) % r for i in range(n)]
    hs, t = [], zt * dinv % r
    for _ in range(h_len):
        hs.append(t)
        t = t * tau % r
    g1 = dict(alpha=[alpha], gamma_abc=[x * ginv % r for x in abc[:p]], beta=[beta], delta=[delta], a=u, b=v, h=hs,
              l=[x * dinv % r for x in abc[p:]])
    g2 = dict(beta=[beta], gamma=[gamma], delta=[delta], b=v)
    trap = dict(tau=tau, alpha=alpha, beta=beta, gamma=gamma, delta=delta, u=u, v=v, w=w, zt=zt)
    return g1, g2, trap


class TrapdoorChecker:
    """Closed-form proofs for many assignments of one circuit (the per-variable constants are prepared once)."""

    def __init__(self, cs, trap):
        self.r, self.p, self.n = cs.modulus, cs.p, cs.n
        r = self.r
        self.u, self.v, self.w = trap["u"], trap["v"], trap["w"]
        self.al, self.be, self.de, self.zt = trap["alpha"], trap["beta"], trap["delta"], trap["zt"]
        self.dinv = pow(self.de, -1, r)
        self.zt_inv = pow(self.zt, -1, r)
        self.abc_aux = [(self.be * a + self.al * b + c) % r for a, b, c in zip(self.u[self.p:], self.v[self.p:], self.w[self.p:])]

    def scalars(self, z, r_rand, s_rand):
        """Discrete logs (a, b, c) of the proof elements w.r.t. the standard generators."""
        r = self.r
        At = sum(map(mul, z, self.u)) % r
        Bt = sum(map(mul, z, self.v)) % r
        Ct = sum(map(mul, z, self.w)) % r
        a_s = (self.al + At + r_rand * self.de) % r
        b_s = (self.be + Bt + s_rand * self.de) % r
        ht = (At * Bt - Ct) * self.zt_inv % r
        c_s = (sum(map(mul, z[self.p:], self.abc_aux)) * self.dinv + ht * self.zt * self.dinv
               + s_rand * a_s + r_rand * b_s - r_rand * s_rand * self.de) % r
        return a_s, b_s, c_s

    @staticmethod
    def bytes_from_scalars(triples) -> list:
        """192-byte compressed proofs from (a, b, c) discrete logs: fixed-base multiplications by the C++ oracle, then the
        ark-serialize compression of the Python oracle."""
        from . import cref
        from .pyref import groth16 as og
        from .pyref.curves import Group
        from .pyref.fields import BLS12_381
        G1, G2 = Group(BLS12_381, 1), Group(BLS12_381, 2)
        g1 = cref.fixed_base(1, [x for t in triples for x in (t[0], t[2])])
        g2 = cref.fixed_base(2, [t[1] for t in triples])
        out = []
        for i in range(len(triples)):
            a = G1.deserialize_uncompressed(g1[192 * i:192 * i + 96])
            c = G1.deserialize_uncompressed(g1[192 * i + 96:192 * i + 192])
            b = G2.deserialize_uncompressed(g2[192 * i:192 * i + 192])
            out.append(og.proof_to_bytes(BLS12_381, (a, b, c)))
        return out

    def proof_bytes(self, z, r_rand, s_rand) -> bytes:
        return self.bytes_from_scalars([self.scalars(z, r_rand, s_rand)])[0]


def trapdoor_proof_scalars(cs, trap, z, r_rand, s_rand):
    return TrapdoorChecker(cs, trap).scalars(z, r_rand, s_rand)
