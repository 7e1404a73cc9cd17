"""TEST INFRASTRUCTURE ONLY — CPU oracle (Python big-int restatement).

Field towers for BLS12-381 and BN254 as used by the reference's arithmetic
back end (arkworks 0.3.0, un-vendored; pinned by `manta-crypto/Cargo.toml:76-87`).
Only `tests/`, `__graft_entry__.smoke()` and `bench.py --impl reference /
cpu_baseline` may import this package; the product path never does.

Conventions restated from ark-ff 0.3:
  * Fq2 = Fq[u]/(u^2 + 1)            (both curves: non-residue -1)
  * Fq6 = Fq2[v]/(v^3 - xi)          xi = 1+u (BLS12-381), 9+u (BN254)
  * Fq12 = Fq6[w]/(w^2 - v)
  * `Ord` on Fq2 compares c1 first, then c0 (used by the y-sign flag).
Elements: Fq = int, Fq2 = (c0, c1), Fq6 = (a0, a1, a2) of Fq2, Fq12 = (b0, b1) of Fq6.
"""
from __future__ import annotations


class CurveParams:
    """Public constants of one pairing-friendly curve."""

    def __init__(self, name, q, r, b, xi, g1, g2, fr_gen, two_adicity, twist_type, x, x_neg):
        self.name = name
        self.q = q            # base field modulus
        self.r = r            # scalar field modulus
        self.b = b            # G1: y^2 = x^3 + b
        self.xi = xi          # Fq6 non-residue (Fq2 element)
        self.g1 = g1          # affine generator (x, y)
        self.g2 = g2          # affine generator ((x0,x1),(y0,y1))
        self.fr_gen = fr_gen  # Fr multiplicative generator (coset shift of ark-poly)
        self.two_adicity = two_adicity
        self.twist_type = twist_type  # 'M' or 'D'
        self.x = x            # curve parameter |x|
        self.x_neg = x_neg
        self.fq_bytes = (q.bit_length() + 7) // 8
        # ark serializes with room for 2 flag bits: ceil((bits + 2) / 8)
        self.fq_ser_bytes = (q.bit_length() + 2 + 7) // 8
        self.fr_bytes = (r.bit_length() + 7) // 8
        self.fq_limbs = (q.bit_length() + 63) // 64
        self.fr_limbs = (r.bit_length() + 63) // 64
        # b' on the twist
        if twist_type == 'M':
            self.b2 = fq2_mul(q, (b, 0), xi)
        else:
            self.b2 = fq2_mul(q, (b, 0), fq2_inv(q, xi))
        # 2^s-th primitive root of unity: gen^((r-1)/2^s)
        self.root_of_unity = pow(fr_gen, (r - 1) >> two_adicity, r)


# ----------------------------------------------------------------------------
# Fq2 arithmetic on tuples (q passed explicitly so both curves share code)
# ----------------------------------------------------------------------------

def fq2_add(q, a, b):
    return ((a[0] + b[0]) % q, (a[1] + b[1]) % q)


def fq2_sub(q, a, b):
    return ((a[0] - b[0]) % q, (a[1] - b[1]) % q)


def fq2_neg(q, a):
    return ((-a[0]) % q, (-a[1]) % q)


def fq2_mul(q, a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % q, (a[0] * b[1] + a[1] * b[0]) % q)


def fq2_sqr(q, a):
    return ((a[0] + a[1]) * (a[0] - a[1]) % q, 2 * a[0] * a[1] % q)


def fq2_inv(q, a):
    n = pow((a[0] * a[0] + a[1] * a[1]) % q, -1, q)
    return (a[0] * n % q, (-a[1]) * n % q)


def fq2_scalar(q, a, k):
    return (a[0] * k % q, a[1] * k % q)


def fq2_conj(q, a):
    return (a[0], (-a[1]) % q)


def fq2_sqrt(q, a):
    """Square root in Fq2 for q = 3 mod 4 (None if `a` is a non-residue)."""
    if a == (0, 0):
        return (0, 0)
    # Algorithm 9 of "Square root computation over even extension fields" (Adj, Rodriguez-Henriquez)
    a1 = fq2_pow(q, a, (q - 3) // 4)
    alpha = fq2_mul(q, fq2_sqr(q, a1), a)
    a0 = fq2_mul(q, fq2_conj(q, alpha), alpha)  # alpha^(q+1)
    if a0 == ((-1) % q, 0):
        return None
    x0 = fq2_mul(q, a1, a)
    if alpha == ((-1) % q, 0):
        x = fq2_mul(q, (0, 1), x0)
    else:
        bb = fq2_pow(q, fq2_add(q, (1, 0), alpha), (q - 1) // 2)
        x = fq2_mul(q, bb, x0)
    return x if fq2_sqr(q, x) == a else None


def fq2_pow(q, a, e):
    res = (1, 0)
    base = a
    while e:
        if e & 1:
            res = fq2_mul(q, res, base)
        base = fq2_sqr(q, base)
        e >>= 1
    return res


class FieldOps:
    """Uniform field interface so the curve code is generic over Fq / Fq2."""

    def __init__(self, q, degree):
        self.q = q
        self.degree = degree
        if degree == 1:
            self.zero, self.one = 0, 1
        else:
            self.zero, self.one = (0, 0), (1, 0)

    def add(self, a, b):
        return (a + b) % self.q if self.degree == 1 else fq2_add(self.q, a, b)

    def sub(self, a, b):
        return (a - b) % self.q if self.degree == 1 else fq2_sub(self.q, a, b)

    def neg(self, a):
        return (-a) % self.q if self.degree == 1 else fq2_neg(self.q, a)

    def mul(self, a, b):
        return a * b % self.q if self.degree == 1 else fq2_mul(self.q, a, b)

    def sqr(self, a):
        return a * a % self.q if self.degree == 1 else fq2_sqr(self.q, a)

    def inv(self, a):
        return pow(a, -1, self.q) if self.degree == 1 else fq2_inv(self.q, a)

    def dbl(self, a):
        return self.add(a, a)

    def small(self, a, k):
        return a * k % self.q if self.degree == 1 else fq2_scalar(self.q, a, k)

    def is_zero(self, a):
        return a == self.zero

    def sqrt(self, a):
        if self.degree == 1:
            s = pow(a, (self.q + 1) // 4, self.q)  # q = 3 mod 4 for both curves
            return s if s * s % self.q == a else None
        return fq2_sqrt(self.q, a)

    def lex_larger(self, y):
        """ark `y > -y` by `Ord` on canonical integers (Fq2: c1 first, then c0). C.8."""
        ny = self.neg(y)
        if self.degree == 1:
            return y > ny
        return (y[1], y[0]) > (ny[1], ny[0])


# ----------------------------------------------------------------------------
# Fq6 / Fq12 (only the pairing KATs need these)
# ----------------------------------------------------------------------------

class Tower:
    def __init__(self, q, xi):
        self.q, self.xi = q, xi
        self.f2 = FieldOps(q, 2)
        self.zero6 = ((0, 0),) * 3
        self.one6 = ((1, 0), (0, 0), (0, 0))
        self.zero12 = (self.zero6, self.zero6)
        self.one12 = (self.one6, self.zero6)

    # Fq6
    def add6(self, a, b):
        return tuple(fq2_add(self.q, x, y) for x, y in zip(a, b))

    def sub6(self, a, b):
        return tuple(fq2_sub(self.q, x, y) for x, y in zip(a, b))

    def neg6(self, a):
        return tuple(fq2_neg(self.q, x) for x in a)

    def mul6(self, a, b):
        q, xi = self.q, self.xi
        m = fq2_mul
        a0, a1, a2 = a
        b0, b1, b2 = b
        c0 = fq2_add(q, m(q, a0, b0), m(q, xi, fq2_add(q, m(q, a1, b2), m(q, a2, b1))))
        c1 = fq2_add(q, fq2_add(q, m(q, a0, b1), m(q, a1, b0)), m(q, xi, m(q, a2, b2)))
        c2 = fq2_add(q, fq2_add(q, m(q, a0, b2), m(q, a1, b1)), m(q, a2, b0))
        return (c0, c1, c2)

    def mul6_by_v(self, a):
        return (fq2_mul(self.q, self.xi, a[2]), a[0], a[1])

    def inv6(self, a):
        q, xi = self.q, self.xi
        a0, a1, a2 = a
        t0 = fq2_sub(q, fq2_sqr(q, a0), fq2_mul(q, xi, fq2_mul(q, a1, a2)))
        t1 = fq2_sub(q, fq2_mul(q, xi, fq2_sqr(q, a2)), fq2_mul(q, a0, a1))
        t2 = fq2_sub(q, fq2_sqr(q, a1), fq2_mul(q, a0, a2))
        d = fq2_add(q, fq2_mul(q, a0, t0),
                    fq2_mul(q, xi, fq2_add(q, fq2_mul(q, a2, t1), fq2_mul(q, a1, t2))))
        di = fq2_inv(q, d)
        return (fq2_mul(q, t0, di), fq2_mul(q, t1, di), fq2_mul(q, t2, di))

    # Fq12
    def mul12(self, a, b):
        a0, a1 = a
        b0, b1 = b
        t0 = self.mul6(a0, b0)
        t1 = self.mul6(a1, b1)
        c0 = self.add6(t0, self.mul6_by_v(t1))
        c1 = self.sub6(self.sub6(self.mul6(self.add6(a0, a1), self.add6(b0, b1)), t0), t1)
        return (c0, c1)

    def sqr12(self, a):
        return self.mul12(a, a)

    def inv12(self, a):
        a0, a1 = a
        d = self.sub6(self.mul6(a0, a0), self.mul6_by_v(self.mul6(a1, a1)))
        di = self.inv6(d)
        return (self.mul6(a0, di), self.neg6(self.mul6(a1, di)))

    def add12(self, a, b):
        return (self.add6(a[0], b[0]), self.add6(a[1], b[1]))

    def sub12(self, a, b):
        return (self.sub6(a[0], b[0]), self.sub6(a[1], b[1]))

    def pow12(self, a, e):
        res = self.one12
        base = a
        while e:
            if e & 1:
                res = self.mul12(res, base)
            base = self.sqr12(base)
            e >>= 1
        return res

    def from_fq(self, x):
        return (((x % self.q, 0), (0, 0), (0, 0)), self.zero6)

    def from_fq2_at(self, x, half, idx):
        """Place Fq2 element `x` at coefficient `idx` of Fq6 half `half`."""
        h = [(0, 0)] * 3
        h[idx] = x
        return (tuple(h), self.zero6) if half == 0 else (self.zero6, tuple(h))


# ----------------------------------------------------------------------------
# Curve constants (public, self-checked in tests)
# ----------------------------------------------------------------------------

BLS12_381 = CurveParams(
    name="bls12_381",
    q=0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab,
    r=0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001,
    b=4,
    xi=(1, 1),
    g1=(3685416753713387016781088315183077757961620795782546409894578378688607592378376318836054947676345821548104185464507,
        1339506544944476473020471379941921221584933875938349620426543736416511423956333506472724655353366534992391756441569),
    g2=((352701069587466618187139116011060144890029952792775240219908644239793785735715026873347600343865175952761926303160,
         3059144344244213709971259814753781636986470325476647558659373206291635324768958432433509563104347017837885763365758),
        (1985150602287291935568054521177171638300868978215655730859378665066344726373823718423869104263333984641494340347905,
         927553665492332455747201965776037880757740193453592970025027978793976877002675564980949289727957565575433344219582)),
    fr_gen=7,
    two_adicity=32,
    twist_type='M',
    x=0xd201000000010000,
    x_neg=True,
)

BN254 = CurveParams(
    name="bn254",
    q=21888242871839275222246405745257275088696311157297823662689037894645226208583,
    r=21888242871839275222246405745257275088548364400416034343698204186575808495617,
    b=3,
    xi=(9, 1),
    g1=(1, 2),
    g2=((10857046999023057135944570762232829481370756359578518086990519993285655852781,
         11559732032986387107991004021392285783925812861821192530917403151452391805634),
        (8495653923123431417604973247489272438418190587263600148770280649306958101930,
         4082367875863433681332203403145435568316851327593401208105741076214120093531)),
    fr_gen=5,
    two_adicity=28,
    twist_type='D',
    x=4965661367192848881,
    x_neg=False,
)

CURVES = {"bls12_381": BLS12_381, "bn254": BN254}
