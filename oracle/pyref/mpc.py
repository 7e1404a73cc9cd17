"""TEST INFRASTRUCTURE ONLY — CPU oracle: the phase-2 `initialize` of the trusted setup, restated from
`manta-trusted-setup/src/groth16/mpc.rs:245-312` (`specialize_to_phase_2`, `add_dummy_constraints`) and `:355-431`
(`initialize`), with ark-poly's `domain.ifft` over curve points (`DomainCoeff`): the same serial radix-2 butterflies as
`poly.Radix2Domain._serial_fft`, group additions in place of field additions and scalar multiplications by the twiddles.
Python big integers; meant for small domains only.
"""
from __future__ import annotations

from .curves import Group
from .poly import Radix2Domain


def group_fft(G: Group, points, dom: Radix2Domain, inverse: bool):
    """ark `domain.fft` / `domain.ifft` on affine points (None = infinity); natural order in and out."""
    n, log_n, r = dom.size, dom.log_size, dom.r
    a = [G.to_jac(P) for P in points[:n]] + [G.jac_identity()] * (n - min(n, len(points)))
    omega = dom.group_gen_inv if inverse else dom.group_gen
    for k in range(n):
        rk = int(format(k, "0%db" % log_n)[::-1], 2) if log_n else 0
        if k < rk:
            a[k], a[rk] = a[rk], a[k]
    m = 1
    for _ in range(log_n):
        w_m = pow(omega, n // (2 * m), r)
        ws = [1] * m
        for j in range(1, m):
            ws[j] = ws[j - 1] * w_m % r
        for k in range(0, n, 2 * m):
            for j in range(m):
                t = G.jac_mul(a[k + j + m], ws[j])
                a[k + j + m] = G.jac_add(a[k + j], G.jac_neg(t))
                a[k + j] = G.jac_add(a[k + j], t)
        m *= 2
    if inverse:
        a = [G.jac_mul(x, dom.size_inv) for x in a]
    return [G.to_affine(x) for x in a]


def initialize(curve, powers, r1cs):
    """mpc.rs:355-431.  powers: dict(tau_powers_g1, tau_powers_g2, alpha_tau_powers_g1, beta_tau_powers_g1, beta_g2) of affine
    points; r1cs: dict(p, w, K, a, b, c).  Returns the proving key dict of pyref.groth16 (gamma = delta = 1)."""
    G1, G2 = Group(curve, 1), Group(curve, 2)
    K, p, n = r1cs["K"], r1cs["p"], r1cs["p"] + r1cs["w"]
    dom = Radix2Domain(curve, K + p)
    degree = dom.size
    tau1 = powers["tau_powers_g1"]
    h_query = [G1.add(tau1[i + degree], G1.neg(tau1[i])) for i in range(degree)]
    tau_lagrange_g1 = group_fft(G1, tau1, dom, True)
    tau_lagrange_g2 = group_fft(G2, powers["tau_powers_g2"], dom, True)
    alpha_lagrange_g1 = group_fft(G1, powers["alpha_tau_powers_g1"], dom, True)
    beta_lagrange_g1 = group_fft(G1, powers["beta_tau_powers_g1"], dom, True)
    a_g1, b_g1, b_g2, ext = [None] * n, [None] * n, [None] * n, [None] * n
    # add_dummy_constraints (:295-312)
    for i in range(p):
        a_g1[i] = tau_lagrange_g1[K + i]
        ext[i] = beta_lagrange_g1[K + i]
    # specialize_to_phase_2 (:245-293)
    for j in range(K):
        for coeff, idx in r1cs["a"][j]:
            a_g1[idx] = G1.add(a_g1[idx], G1.mul(tau_lagrange_g1[j], coeff))
            ext[idx] = G1.add(ext[idx], G1.mul(beta_lagrange_g1[j], coeff))
        for coeff, idx in r1cs["b"][j]:
            b_g1[idx] = G1.add(b_g1[idx], G1.mul(tau_lagrange_g1[j], coeff))
            b_g2[idx] = G2.add(b_g2[idx], G2.mul(tau_lagrange_g2[j], coeff))
            ext[idx] = G1.add(ext[idx], G1.mul(alpha_lagrange_g1[j], coeff))
        for coeff, idx in r1cs["c"][j]:
            ext[idx] = G1.add(ext[idx], G1.mul(tau_lagrange_g1[j], coeff))
    return {
        "vk": {"alpha_g1": powers["alpha_tau_powers_g1"][0], "beta_g2": powers["beta_g2"], "gamma_g2": curve.g2, "delta_g2": curve.g2,
               "gamma_abc_g1": ext[:p]},
        "beta_g1": powers["beta_tau_powers_g1"][0], "delta_g1": curve.g1,
        "a_query": a_g1, "b_g1_query": b_g1, "b_g2_query": b_g2, "h_query": h_query, "l_query": ext[p:],
    }
