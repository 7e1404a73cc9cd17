"""Shared test helpers (CPU oracle side).  The oracle is the checker; the product path never imports it."""
from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import cref  # noqa: E402
from oracle.pyref.fields import BLS12_381  # noqa: E402
import manta_rs_b200  # noqa: E402,F401
from manta_rs_b200 import keygen as _keygen  # noqa: E402  (host-side QAP arithmetic only; no GPU call here)


def oracle_keygen(cs, trapdoor, h_len=None):
    """Known-trapdoor proving key built entirely on the CPU: QAP evaluation with Python integers and fixed-base
    multiplications by the C++ oracle.  Returns (`ProvingContext` bytes, trapdoor dict)."""
    tau, alpha, beta, gamma, delta = trapdoor
    r = cs.modulus
    u, v, w, zt = _keygen.qap_at_tau(cs, tau)
    n, p, m = cs.n, cs.p, cs.m
    h_len = m - 1 if h_len is None else h_len
    ginv, dinv = pow(gamma, -1, r), pow(delta, -1, r)
    abc = [(beta * u[i] + alpha * v[i] + w[i]) % r for i in range(n)]
    hs, t = [], zt * dinv % r
    for _ in range(h_len):
        hs.append(t)
        t = t * tau % r
    g1 = cref.fixed_base(1, [alpha] + [x * ginv % r for x in abc[:p]] + [beta, delta] + u + v + hs + [x * dinv % r for x in abc[p:]])
    g2 = cref.fixed_base(2, [beta, gamma, delta] + v)
    pos = [0]

    def take(k):
        s = g1[pos[0] * 96:(pos[0] + k) * 96]
        pos[0] += k
        return s

    def vec(data, k):
        return k.to_bytes(8, "little") + data

    alpha_g1, gamma_abc, beta_g1, delta_g1 = take(1), take(p), take(1), take(1)
    a_q, b1_q, h_q, l_q = take(n), take(n), take(h_len), take(n - p)
    pk = (alpha_g1 + g2[:192] + g2[192:384] + g2[384:576] + vec(gamma_abc, p) + beta_g1 + delta_g1 + vec(a_q, n)
          + vec(b1_q, n) + vec(g2[576:], n) + vec(h_q, h_len) + vec(l_q, n - p))
    trap = dict(tau=tau, alpha=alpha, beta=beta, gamma=gamma, delta=delta, u=u, v=v, w=w, zt=zt)
    return pk, trap


def trapdoor_proof_bytes(cs, trap, z, r_rand, s_rand) -> bytes:
    """Proof bytes from the closed form (three oracle scalar multiplications + oracle compression)."""
    from oracle.pyref import groth16 as og
    from oracle.pyref.curves import Group
    a_s, b_s, c_s = _keygen.trapdoor_proof_scalars(cs, trap, z, r_rand, s_rand)
    G1, G2 = Group(BLS12_381, 1), Group(BLS12_381, 2)
    a = G1.deserialize_uncompressed(cref.fixed_base(1, [a_s]))
    b = G2.deserialize_uncompressed(cref.fixed_base(2, [b_s]))
    c = G1.deserialize_uncompressed(cref.fixed_base(1, [c_s]))
    return og.proof_to_bytes(BLS12_381, (a, b, c))
