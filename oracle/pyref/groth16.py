"""TEST INFRASTRUCTURE ONLY — CPU oracle: ark-groth16 0.3 prover path restated in Python big ints.

Follows SURVEY.md §3.1 / Appendix C.3-C.7 (upstream ark-groth16 0.3.0 `prover.rs`,
`r1cs_to_qap.rs`, `generator.rs`; not vendored — the reference reaches them at
`manta-crypto/src/arkworks/groth16.rs:581,597`; the in-repo re-derivation of the key
structure is `manta-trusted-setup/src/groth16/mpc.rs:245-312,355-431`).

R1CS container (C.3): dict(p=num_instance incl. the constant 1, w=num_witness, K=num_constraints,
a/b/c = row-major sparse rows [(coeff, column)], columns: instance first then witness).
"""
from __future__ import annotations

from .curves import Group, msm_pippenger
from .poly import Radix2Domain


def eval_rows(rows, z, r):
    return [sum(c * z[i] for c, i in row) % r for row in rows]


def witness_map(curve, r1cs, z):
    """C.4 — returns (h coefficients [m], domain)."""
    r = curve.r
    K, p = r1cs["K"], r1cs["p"]
    dom = Radix2Domain(curve, K + p)
    m = dom.size
    a = eval_rows(r1cs["a"], z, r) + [0] * (m - K)
    b = eval_rows(r1cs["b"], z, r) + [0] * (m - K)
    for j in range(p):
        a[K + j] = z[j]
    a = dom.coset_fft(dom.ifft(a))
    b = dom.coset_fft(dom.ifft(b))
    ab = [x * y % r for x, y in zip(a, b)]
    c = eval_rows(r1cs["c"], z, r) + [0] * (m - K)
    c = dom.coset_fft(dom.ifft(c))
    zinv = pow(dom.vanishing_on_coset(), -1, r)
    ab = [(x - y) * zinv % r for x, y in zip(ab, c)]
    return dom.coset_ifft(ab), dom


def qap_at_tau(curve, r1cs, tau):
    """u_i(tau), v_i(tau), w_i(tau) for every variable i, plus (Z(tau), domain). C.7."""
    r = curve.r
    K, p, n = r1cs["K"], r1cs["p"], r1cs["p"] + r1cs["w"]
    dom = Radix2Domain(curve, K + p)
    L = dom.lagrange_at(tau)
    u, v, w = [0] * n, [0] * n, [0] * n
    for i in range(p):
        u[i] = L[K + i]
    for j in range(K):
        for coeff, col in r1cs["a"][j]:
            u[col] = (u[col] + L[j] * coeff) % r
        for coeff, col in r1cs["b"][j]:
            v[col] = (v[col] + L[j] * coeff) % r
        for coeff, col in r1cs["c"][j]:
            w[col] = (w[col] + L[j] * coeff) % r
    zt = (pow(tau, dom.size, r) - 1) % r
    return u, v, w, zt, dom


def setup_trapdoor(curve, r1cs, tau, alpha, beta, gamma, delta, g1_gen=None, g2_gen=None, h_len=None):
    """Known-toxic-waste Groth16 key (C.7).  `h_len` = m-1 (ark generator, default) or m (MPC keys,
    `manta-trusted-setup/src/groth16/mpc.rs:371-377`)."""
    r = curve.r
    G1, G2 = Group(curve, 1), Group(curve, 2)
    g1_gen = g1_gen or curve.g1
    g2_gen = g2_gen or curve.g2
    p, n = r1cs["p"], r1cs["p"] + r1cs["w"]
    u, v, w, zt, dom = qap_at_tau(curve, r1cs, tau)
    m = dom.size
    h_len = m - 1 if h_len is None else h_len
    fb1 = G1.fixed_base(g1_gen, r.bit_length())
    fb2 = G2.fixed_base(g2_gen, r.bit_length())

    def g1s(scalars):
        return G1.batch_to_affine([fb1.mul_jac(k % r) for k in scalars])

    def g2s(scalars):
        return G2.batch_to_affine([fb2.mul_jac(k % r) for k in scalars])

    ginv, dinv = pow(gamma, -1, r), pow(delta, -1, r)
    abc = [(beta * u[i] + alpha * v[i] + w[i]) % r for i in range(n)]
    pk = {
        "vk": {
            "alpha_g1": g1s([alpha])[0],
            "beta_g2": g2s([beta])[0],
            "gamma_g2": g2s([gamma])[0],
            "delta_g2": g2s([delta])[0],
            "gamma_abc_g1": g1s([abc[i] * ginv for i in range(p)]),
        },
        "beta_g1": g1s([beta])[0],
        "delta_g1": g1s([delta])[0],
        "a_query": g1s(u),
        "b_g1_query": g1s(v),
        "b_g2_query": g2s(v),
        "h_query": g1s([pow(tau, k, r) * zt % r * dinv for k in range(h_len)]),
        "l_query": g1s([abc[i] * dinv for i in range(p, n)]),
    }
    trap = dict(tau=tau, alpha=alpha, beta=beta, gamma=gamma, delta=delta, u=u, v=v, w=w, zt=zt,
                g1_gen=g1_gen, g2_gen=g2_gen)
    return pk, trap


def create_proof(curve, pk, r1cs, z, r_rand, s_rand, msm=msm_pippenger):
    """C.6 — returns affine (A in G1, B in G2, C in G1)."""
    G1, G2 = Group(curve, 1), Group(curve, 2)
    p = r1cs["p"]
    h, _ = witness_map(curve, r1cs, z)
    assignment = z[1:]
    aux = z[p:]
    h_acc = msm(G1, pk["h_query"], h)
    l_acc = msm(G1, pk["l_query"], aux)

    def calculate_coeff(G, initial, query, vk_param):
        acc = msm(G, query[1:], assignment)
        res = G.jac_add_mixed(initial, query[0])
        res = G.jac_add(res, acc)
        return G.jac_add_mixed(res, vk_param)

    r_delta = G1.jac_mul(G1.to_jac(pk["delta_g1"]), r_rand)
    g_a = calculate_coeff(G1, r_delta, pk["a_query"], pk["vk"]["alpha_g1"])
    s_delta1 = G1.jac_mul(G1.to_jac(pk["delta_g1"]), s_rand)
    g1_b = calculate_coeff(G1, s_delta1, pk["b_g1_query"], pk["beta_g1"]) if r_rand != 0 else G1.jac_identity()
    s_delta2 = G2.jac_mul(G2.to_jac(pk["vk"]["delta_g2"]), s_rand)
    g2_b = calculate_coeff(G2, s_delta2, pk["b_g2_query"], pk["vk"]["beta_g2"])
    g_c = G1.jac_mul(g_a, s_rand)
    g_c = G1.jac_add(g_c, G1.jac_mul(g1_b, r_rand))
    rs_delta = G1.jac_mul(G1.jac_mul(G1.to_jac(pk["delta_g1"]), r_rand), s_rand)
    g_c = G1.jac_add(g_c, G1.jac_neg(rs_delta))
    g_c = G1.jac_add(g_c, l_acc)
    g_c = G1.jac_add(g_c, h_acc)
    return G1.to_affine(g_a), G2.to_affine(g2_b), G1.to_affine(g_c)


def proof_to_bytes(curve, proof) -> bytes:
    """`proof_as_bytes` (`manta-crypto/src/arkworks/groth16.rs:184-195`): compressed a ‖ b ‖ c."""
    G1, G2 = Group(curve, 1), Group(curve, 2)
    return G1.compress(proof[0]) + G2.compress(proof[1]) + G1.compress(proof[2])


def proof_from_bytes(curve, data: bytes):
    G1, G2 = Group(curve, 1), Group(curve, 2)
    n1, n2 = G1.coord_bytes, G2.coord_bytes
    return (G1.decompress(data[:n1]), G2.decompress(data[n1:n1 + n2]), G1.decompress(data[n1 + n2:n1 + n2 + n1]))


def trapdoor_proof(curve, r1cs, trap, z, r_rand, s_rand):
    """Closed form of the proof from the toxic waste (C.7) — Fr arithmetic + three scalar muls only.
    Independent of every MSM / NTT code path."""
    r = curve.r
    G1, G2 = Group(curve, 1), Group(curve, 2)
    p = r1cs["p"]
    u, v, w = trap["u"], trap["v"], trap["w"]
    al, be, de, tau, zt = trap["alpha"], trap["beta"], trap["delta"], trap["tau"], trap["zt"]
    n = len(z)
    a_s = (al + sum(z[i] * u[i] for i in range(n)) + r_rand * de) % r
    b_s = (be + sum(z[i] * v[i] for i in range(n)) + s_rand * de) % r
    # h(tau) = (A(tau) B(tau) - C(tau)) / Z(tau) with A(tau) = sum z_i u_i etc.
    At = sum(z[i] * u[i] for i in range(n)) % r
    Bt = sum(z[i] * v[i] for i in range(n)) % r
    Ct = sum(z[i] * w[i] for i in range(n)) % r
    ht = (At * Bt - Ct) * pow(zt, -1, r) % r
    dinv = pow(de, -1, r)
    c_s = (sum(z[i] * (be * u[i] + al * v[i] + w[i]) for i in range(p, n)) * dinv
           + ht * zt * dinv + s_rand * a_s + r_rand * b_s - r_rand * s_rand * de) % r
    return (G1.mul(trap["g1_gen"], a_s), G2.mul(trap["g2_gen"], b_s), G1.mul(trap["g1_gen"], c_s))


# ----------------------------------------------------------------------------
# ProvingContext file format (`groth16.rs:268-303`, `serialize_unchecked` = uncompressed, C.8)
# ----------------------------------------------------------------------------

def pk_to_bytes(curve, pk) -> bytes:
    G1, G2 = Group(curve, 1), Group(curve, 2)
    out = bytearray()

    def vec(G, pts):
        out.extend(len(pts).to_bytes(8, "little"))
        for P in pts:
            out.extend(G.serialize_uncompressed(P))

    vk = pk["vk"]
    out += G1.serialize_uncompressed(vk["alpha_g1"])
    out += G2.serialize_uncompressed(vk["beta_g2"])
    out += G2.serialize_uncompressed(vk["gamma_g2"])
    out += G2.serialize_uncompressed(vk["delta_g2"])
    vec(G1, vk["gamma_abc_g1"])
    out += G1.serialize_uncompressed(pk["beta_g1"])
    out += G1.serialize_uncompressed(pk["delta_g1"])
    vec(G1, pk["a_query"])
    vec(G1, pk["b_g1_query"])
    vec(G2, pk["b_g2_query"])
    vec(G1, pk["h_query"])
    vec(G1, pk["l_query"])
    return bytes(out)


def pk_from_bytes(curve, data: bytes):
    G1, G2 = Group(curve, 1), Group(curve, 2)
    pos = 0

    def pt(G):
        nonlocal pos
        n = 2 * G.coord_bytes
        P = G.deserialize_uncompressed(data[pos:pos + n])
        pos += n
        return P

    def vec(G):
        nonlocal pos
        cnt = int.from_bytes(data[pos:pos + 8], "little")
        pos += 8
        return [pt(G) for _ in range(cnt)]

    vk = {"alpha_g1": pt(G1), "beta_g2": pt(G2), "gamma_g2": pt(G2), "delta_g2": pt(G2)}
    vk["gamma_abc_g1"] = vec(G1)
    pk = {"vk": vk, "beta_g1": pt(G1), "delta_g1": pt(G1)}
    pk["a_query"] = vec(G1)
    pk["b_g1_query"] = vec(G1)
    pk["b_g2_query"] = vec(G2)
    pk["h_query"] = vec(G1)
    pk["l_query"] = vec(G1)
    assert pos == len(data)
    return pk
