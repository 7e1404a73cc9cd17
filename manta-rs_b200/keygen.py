"""Synthetic proving keys for the benchmark workloads.

The reference generates its benchmark keys with `Groth16::compile` (groth16.rs:570-586 ->
ark `circuit_specific_setup`) from a seeded rng (`manta-pay/src/parameters.rs:56-106`).  Here the QAP
evaluations at tau are plain Fr arithmetic on the host (Python integers) and the ~5n + m fixed-base
scalar multiplications run on the GPU through `mp_fixed_base_g1/g2`.  Output: the reference's
`ProvingContext` byte format (groth16.rs:290-303) plus the trapdoor for closed-form proof checks.
"""
from __future__ import annotations

import ctypes

from . import _native as nat


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


def _qap_at_tau(cs, tau, root_of_unity_2_32=None):
    """u_i(tau), v_i(tau), w_i(tau) for all variables, Z(tau), m  (SURVEY.md C.7)."""
    r = cs.modulus
    m, log_m = cs.m, cs.log_m
    root = pow(7, (r - 1) >> 32, r) if root_of_unity_2_32 is None else root_of_unity_2_32
    omega = pow(root, 1 << (32 - log_m), r)
    zt = (pow(tau, m, r) - 1) % r
    assert zt != 0
    # L_j(tau) = Z(tau)/m * w^j / (tau - w^j)
    ws, wj = [], 1
    for _ in range(m):
        ws.append(wj)
        wj = wj * omega % r
    invs = _batch_inv([(tau - x) % r for x in ws], r)
    zm = zt * pow(m, -1, r) % r
    L = [zm * x % r * y % r for x, y in zip(ws, invs)]
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


def _fixed_base(group: int, scalars, device: int) -> bytes:
    pb = nat.G1_BYTES if group == 1 else nat.G2_BYTES
    out = ctypes.create_string_buffer(max(len(scalars), 1) * pb)
    fn = nat.lib().mp_fixed_base_g1 if group == 1 else nat.lib().mp_fixed_base_g2
    nat.check(fn(device, nat.pack_scalars(scalars), len(scalars), out))
    return out.raw[: len(scalars) * pb]


def generate(cs, trapdoor, device: int = 0, h_len=None):
    """Returns the proving-key bytes in `ProvingContext` format."""
    tau, alpha, beta, gamma, delta = trapdoor
    r = cs.modulus
    u, v, w, zt = _qap_at_tau(cs, tau)
    n, p, m = cs.n, cs.p, cs.m
    h_len = m - 1 if h_len is None else h_len
    ginv, dinv = pow(gamma, -1, r), pow(delta, -1, r)
    abc = [(beta * u[i] + alpha * v[i] + w[i]) % r for i in range(n)]
    hs, t = [], zt * dinv % r
    for _ in range(h_len):
        hs.append(t)
        t = t * tau % r
    g1_scalars = ([alpha] + [x * ginv % r for x in abc[:p]] + [beta, delta] + u + v + hs
                  + [x * dinv % r for x in abc[p:]])
    g1 = _fixed_base(1, g1_scalars, device)
    g2 = _fixed_base(2, [beta, gamma, delta] + v, device)
    P1, P2 = nat.G1_BYTES, nat.G2_BYTES
    pos = [0]

    def take1(k):
        s = g1[pos[0] * P1:(pos[0] + k) * P1]
        pos[0] += k
        return s

    def vec(data, k):
        return k.to_bytes(8, "little") + data

    alpha_g1 = take1(1)
    gamma_abc = take1(p)
    beta_g1 = take1(1)
    delta_g1 = take1(1)
    a_q = take1(n)
    b1_q = take1(n)
    h_q = take1(h_len)
    l_q = take1(n - p)
    beta_g2, gamma_g2, delta_g2 = g2[:P2], g2[P2:2 * P2], g2[2 * P2:3 * P2]
    b2_q = g2[3 * P2:]
    pk = (alpha_g1 + beta_g2 + gamma_g2 + delta_g2 + vec(gamma_abc, p) + beta_g1 + delta_g1 + vec(a_q, n)
          + vec(b1_q, n) + vec(b2_q, n) + vec(h_q, h_len) + vec(l_q, n - p))
    return pk
