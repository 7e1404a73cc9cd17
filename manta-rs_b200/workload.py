"""Synthetic satisfiable R1CS instances at the reference's circuit shapes (host-side input generation).

The reference builds its constraint systems with ark-r1cs-std gadgets
(`manta-accounting/src/transfer/mod.rs:667-673,889-915`) — out of scope.  The
benchmark workloads are synthetic systems with the reference's sizes (derived from
`manta-parameters/data/pay/proving/*.lfs` and `data/pay/verifying/*.dat`) built by the recipe of
SURVEY.md §8d config 1:

  * p instance variables (z_0 = 1), w witnesses, K = w constraints, domain m = 2^ceil(log2(K + p));
  * constraint i defines witness p+i:  <A_i, z> * <B_i, z> = z_{p+i}, A_i and B_i = 3 random earlier
    columns each with uniform Fr coefficients, C_i = e_{p+i};
  * distribution "R" overrides ~10 % of the witnesses with booleans (row b * (1 - b) = 0) and ~1 % with
    values < 2^128 (row v * 1 = v), the estimate of SURVEY.md §8d config 2; "U" is all-uniform.

Only Python integers mod r are used here: this is input generation, not the proving path.
"""
from __future__ import annotations

import random
from dataclasses import dataclass, field

FR_BLS12_381 = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
SEED_MANTA = 0x4D414E5441  # "MANTA"

# name -> (n variables incl. "1", p public incl. "1", w witnesses, log2 m)   SURVEY.md §8
SHAPES = {
    "to_private": (8253, 13, 8240, 14),
    "private_transfer": (35175, 27, 35148, 16),
    "to_public": (27945, 19, 27926, 15),
}

KIND_MUL, KIND_BOOL, KIND_SMALL = 0, 1, 2


@dataclass
class R1CS:
    """Row-major sparse A, B, C over Fr (ark `to_matrices()` layout)."""
    modulus: int
    p: int
    w: int
    K: int
    a: list = field(default_factory=list)
    b: list = field(default_factory=list)
    c: list = field(default_factory=list)
    kinds: list = field(default_factory=list)

    @property
    def n(self):
        return self.p + self.w

    @property
    def log_m(self):
        return max((self.K + self.p - 1).bit_length(), 1)

    @property
    def m(self):
        return 1 << self.log_m

    def as_dict(self):
        return {"p": self.p, "w": self.w, "K": self.K, "a": self.a, "b": self.b, "c": self.c}


def make_r1cs(p: int, w: int, seed: int = SEED_MANTA, dist: str = "U", modulus: int = FR_BLS12_381,
              fan_in: int = 3) -> R1CS:
    rng = random.Random(seed)
    cs = R1CS(modulus=modulus, p=p, w=w, K=w)
    for i in range(w):
        col = p + i
        kind = KIND_MUL
        if dist == "R":
            u = rng.random()
            kind = KIND_BOOL if u < 0.10 else (KIND_SMALL if u < 0.11 else KIND_MUL)
        cs.kinds.append(kind)
        if kind == KIND_MUL:
            k = min(fan_in, col)
            ca = sorted(rng.sample(range(col), k))
            cb = sorted(rng.sample(range(col), k))
            cs.a.append([(rng.randrange(1, modulus), j) for j in ca])
            cs.b.append([(rng.randrange(1, modulus), j) for j in cb])
            cs.c.append([(1, col)])
        elif kind == KIND_BOOL:
            cs.a.append([(1, col)])
            cs.b.append([(1, 0), (modulus - 1, col)])
            cs.c.append([])
        else:
            cs.a.append([(1, col)])
            cs.b.append([(1, 0)])
            cs.c.append([(1, col)])
    return cs


def make_shape(name: str, seed: int = SEED_MANTA, dist: str = "U") -> R1CS:
    n, p, w, log_m = SHAPES[name]
    cs = make_r1cs(p, w, seed=seed, dist=dist)
    assert cs.n == n and cs.log_m == log_m
    return cs


def make_assignment(cs: R1CS, seed: int) -> list:
    """Full assignment z (canonical integers), z_0 = 1, satisfying every constraint."""
    r = cs.modulus
    rng = random.Random((seed << 20) ^ 0x7A)
    z = [1] + [rng.randrange(r) for _ in range(cs.p - 1)] + [0] * cs.w
    for i in range(cs.w):
        kind = cs.kinds[i]
        if kind == KIND_MUL:
            sa = 0
            for coeff, j in cs.a[i]:
                sa += coeff * z[j]
            sb = 0
            for coeff, j in cs.b[i]:
                sb += coeff * z[j]
            z[cs.p + i] = (sa % r) * (sb % r) % r
        elif kind == KIND_BOOL:
            z[cs.p + i] = rng.getrandbits(1)
        else:
            z[cs.p + i] = rng.getrandbits(128)
    return z


def is_satisfied(cs: R1CS, z) -> bool:
    r = cs.modulus
    for ra, rb, rc in zip(cs.a, cs.b, cs.c):
        va = sum(c * z[j] for c, j in ra) % r
        vb = sum(c * z[j] for c, j in rb) % r
        vc = sum(c * z[j] for c, j in rc) % r
        if va * vb % r != vc:
            return False
    return True


def sample_trapdoor(seed: int, modulus: int = FR_BLS12_381):
    """(tau, alpha, beta, gamma, delta) from the seeded stream (all non-zero)."""
    rng = random.Random((seed << 8) ^ 0x51)
    return tuple(rng.randrange(2, modulus) for _ in range(5))
