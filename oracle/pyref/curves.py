"""TEST INFRASTRUCTURE ONLY — CPU oracle: short-Weierstrass groups + ark-serialize 0.3 formats.

Restates (upstream ark-ec 0.3.0
`models/short_weierstrass_jacobian.rs`, ark-serialize 0.3.0 — not vendored in
/root/reference; call sites `manta-crypto/src/arkworks/groth16.rs:184-195,268-303`):
  * Jacobian add / mixed add / double (a = 0 curves)
  * compressed:   x LE with flags in the top bits of the LAST byte
                  (0x80 = y is the lexicographically larger of {y,-y}; 0x40 = infinity)
  * uncompressed: x ‖ y, infinity flag (0x40) on the last byte of y
  * Fq2 is written c0 ‖ c1, flags live on c1's last byte.
Points: affine = (x, y) or None for infinity; Jacobian = (X, Y, Z) with Z == 0 for infinity.
"""
from __future__ import annotations

from .fields import FieldOps, CurveParams


class Group:
    def __init__(self, curve: CurveParams, which: int):
        self.curve = curve
        self.which = which
        self.F = FieldOps(curve.q, 1 if which == 1 else 2)
        self.b = curve.b if which == 1 else curve.b2
        self.gen = curve.g1 if which == 1 else curve.g2
        self.coord_bytes = curve.fq_ser_bytes * (1 if which == 1 else 2)

    # ---- predicates -------------------------------------------------------
    def on_curve(self, P):
        if P is None:
            return True
        F = self.F
        x, y = P
        return F.sqr(y) == F.add(F.mul(F.sqr(x), x), self.b)

    # ---- Jacobian ---------------------------------------------------------
    def jac_identity(self):
        return (self.F.one, self.F.one, self.F.zero)

    def to_jac(self, P):
        return self.jac_identity() if P is None else (P[0], P[1], self.F.one)

    def to_affine(self, J):
        F = self.F
        X, Y, Z = J
        if F.is_zero(Z):
            return None
        zi = F.inv(Z)
        zi2 = F.sqr(zi)
        return (F.mul(X, zi2), F.mul(Y, F.mul(zi2, zi)))

    def jac_double(self, J):
        """dbl-2009-l (a = 0): 2M + 5S — what ark's `double_in_place` uses."""
        F = self.F
        X, Y, Z = J
        if F.is_zero(Z):
            return J
        A = F.sqr(X)
        B = F.sqr(Y)
        C = F.sqr(B)
        D = F.dbl(F.sub(F.sub(F.sqr(F.add(X, B)), A), C))
        E = F.add(F.dbl(A), A)
        Fv = F.sqr(E)
        Z3 = F.dbl(F.mul(Y, Z))
        X3 = F.sub(Fv, F.dbl(D))
        Y3 = F.sub(F.mul(E, F.sub(D, X3)), F.small(C, 8))
        return (X3, Y3, Z3)

    def jac_add(self, J1, J2):
        """add-2007-bl: 11M + 5S."""
        F = self.F
        if F.is_zero(J1[2]):
            return J2
        if F.is_zero(J2[2]):
            return J1
        X1, Y1, Z1 = J1
        X2, Y2, Z2 = J2
        Z1Z1 = F.sqr(Z1)
        Z2Z2 = F.sqr(Z2)
        U1 = F.mul(X1, Z2Z2)
        U2 = F.mul(X2, Z1Z1)
        S1 = F.mul(F.mul(Y1, Z2), Z2Z2)
        S2 = F.mul(F.mul(Y2, Z1), Z1Z1)
        if U1 == U2:
            if S1 == S2:
                return self.jac_double(J1)
            return self.jac_identity()
        H = F.sub(U2, U1)
        I = F.sqr(F.dbl(H))
        Jv = F.mul(H, I)
        r = F.dbl(F.sub(S2, S1))
        V = F.mul(U1, I)
        X3 = F.sub(F.sub(F.sqr(r), Jv), F.dbl(V))
        Y3 = F.sub(F.mul(r, F.sub(V, X3)), F.dbl(F.mul(S1, Jv)))
        Z3 = F.mul(F.sub(F.sub(F.sqr(F.add(Z1, Z2)), Z1Z1), Z2Z2), H)
        return (X3, Y3, Z3)

    def jac_add_mixed(self, J, P):
        """madd-2007-bl: 7M + 4S — ark's `add_assign_mixed` (the 11 of SURVEY §8d)."""
        F = self.F
        if P is None:
            return J
        if F.is_zero(J[2]):
            return (P[0], P[1], F.one)
        X1, Y1, Z1 = J
        x2, y2 = P
        Z1Z1 = F.sqr(Z1)
        U2 = F.mul(x2, Z1Z1)
        S2 = F.mul(F.mul(y2, Z1), Z1Z1)
        if X1 == U2:
            if Y1 == S2:
                return self.jac_double(J)
            return self.jac_identity()
        H = F.sub(U2, X1)
        HH = F.sqr(H)
        I = F.small(HH, 4)
        Jv = F.mul(H, I)
        r = F.dbl(F.sub(S2, Y1))
        V = F.mul(X1, I)
        X3 = F.sub(F.sub(F.sqr(r), Jv), F.dbl(V))
        Y3 = F.sub(F.mul(r, F.sub(V, X3)), F.dbl(F.mul(Y1, Jv)))
        Z3 = F.sub(F.sub(F.sqr(F.add(Z1, H)), Z1Z1), HH)
        return (X3, Y3, Z3)

    def jac_neg(self, J):
        return (J[0], self.F.neg(J[1]), J[2])

    def neg(self, P):
        return None if P is None else (P[0], self.F.neg(P[1]))

    def add(self, P, Q):
        return self.to_affine(self.jac_add_mixed(self.to_jac(P), Q))

    def jac_mul(self, J, k):
        """Left-to-right double-and-add (k is a plain non-negative integer)."""
        acc = self.jac_identity()
        for bit in bin(k)[2:] if k else "":
            acc = self.jac_double(acc)
            if bit == "1":
                acc = self.jac_add(acc, J)
        return acc

    def mul(self, P, k):
        if P is None or k == 0:
            return None
        acc = self.jac_identity()
        for bit in bin(k)[2:]:
            acc = self.jac_double(acc)
            if bit == "1":
                acc = self.jac_add_mixed(acc, P)
        return self.to_affine(acc)

    def batch_to_affine(self, Js):
        """Montgomery batch inversion (ark `batch_normalization`)."""
        F = self.F
        prods, acc = [], F.one
        for J in Js:
            if not F.is_zero(J[2]):
                acc = F.mul(acc, J[2])
            prods.append(acc)
        inv = F.inv(acc) if not F.is_zero(acc) else acc
        out = [None] * len(Js)
        for i in range(len(Js) - 1, -1, -1):
            J = Js[i]
            if F.is_zero(J[2]):
                continue
            prev = F.one
            for j in range(i - 1, -1, -1):
                if not F.is_zero(Js[j][2]):
                    prev = prods[j]
                    break
            zi = F.mul(inv, prev)
            inv = F.mul(inv, J[2])
            zi2 = F.sqr(zi)
            out[i] = (F.mul(J[0], zi2), F.mul(J[1], F.mul(zi2, zi)))
        return out

    class FixedBase:
        """Windowed fixed-base table (keygen helper; not on the prove path)."""

        def __init__(self, group, P, bits, w=8):
            self.g, self.w = group, w
            self.nwin = (bits + w - 1) // w
            self.table = []
            base = group.to_jac(P)
            for _ in range(self.nwin):
                row_j = [group.jac_identity()]
                for _ in range((1 << w) - 1):
                    row_j.append(group.jac_add(row_j[-1], base))
                self.table.append(group.batch_to_affine(row_j))
                for _ in range(w):
                    base = group.jac_double(base)

        def mul_jac(self, k):
            g = self.g
            acc = g.jac_identity()
            for i in range(self.nwin):
                d = (k >> (i * self.w)) & ((1 << self.w) - 1)
                if d:
                    acc = g.jac_add_mixed(acc, self.table[i][d])
            return acc

    def fixed_base(self, P, bits, w=8):
        return Group.FixedBase(self, P, bits, w)

    # ---- ark-serialize ----------------------------------------------------
    def _fe_to_bytes(self, x):
        n = self.curve.fq_ser_bytes
        if self.which == 1:
            return bytearray(x.to_bytes(n, "little"))
        return bytearray(x[0].to_bytes(n, "little") + x[1].to_bytes(n, "little"))

    def _fe_from_bytes(self, b):
        """Returns (element, flags) with the two flag bits of the last byte stripped."""
        n = self.curve.fq_ser_bytes
        flags = b[-1] & 0xC0
        bb = bytearray(b)
        bb[-1] &= 0x3F
        if self.which == 1:
            v = int.from_bytes(bb[:n], "little")
            assert v < self.curve.q
            return v, flags
        c0 = int.from_bytes(bb[:n], "little")
        c1 = int.from_bytes(bb[n:2 * n], "little")
        assert c0 < self.curve.q and c1 < self.curve.q
        return (c0, c1), flags

    def compress(self, P) -> bytes:
        if P is None:
            out = bytearray(self.coord_bytes)
            out[-1] |= 0x40
            return bytes(out)
        out = self._fe_to_bytes(P[0])
        if self.F.lex_larger(P[1]):
            out[-1] |= 0x80
        return bytes(out)

    def decompress(self, b: bytes):
        x, flags = self._fe_from_bytes(b[: self.coord_bytes])
        if flags & 0x40:
            return None
        F = self.F
        y = F.sqrt(F.add(F.mul(F.sqr(x), x), self.b))
        assert y is not None, "x is not on the curve"
        if F.lex_larger(y) != bool(flags & 0x80):
            y = F.neg(y)
        return (x, y)

    def serialize_uncompressed(self, P) -> bytes:
        if P is None:
            out = bytearray(2 * self.coord_bytes)
            out[-1] |= 0x40
            return bytes(out)
        return bytes(self._fe_to_bytes(P[0]) + self._fe_to_bytes(P[1]))

    def deserialize_uncompressed(self, b: bytes):
        cb = self.coord_bytes
        x, _ = self._fe_from_bytes(b[:cb])
        y, flags = self._fe_from_bytes(b[cb:2 * cb])
        if flags & 0x40:
            return None
        return (x, y)


def msm_naive(group: Group, bases, scalars):
    acc = group.jac_identity()
    for P, k in zip(bases, scalars):
        if k and P is not None:
            acc = group.jac_add(acc, group.jac_mul(group.to_jac(P), k))
    return acc


def msm_pippenger(group: Group, bases, scalars, c=None):
    """ark-ec 0.3 `VariableBaseMSM::multi_scalar_mul` restated:
    window c = 3 if n < 32 else ln_without_floats(n) + 2 where
    ln_without_floats(n) = ceil(log2 n) * 69 / 100; zero scalars dropped; unit scalars
    added directly in window 0; 2^c - 1 Jacobian buckets per window via mixed adds;
    running-sum reduction; Horner combine from the top window with c doublings."""
    size = min(len(bases), len(scalars))
    bases, scalars = bases[:size], scalars[:size]
    if c is None:
        c = 3 if size < 32 else ((size - 1).bit_length() * 69 // 100) + 2
    num_bits = group.curve.r.bit_length()
    pairs = [(b, s) for b, s in zip(bases, scalars) if s != 0]
    window_sums = []
    for w_start in range(0, num_bits, c):
        res = group.jac_identity()
        buckets = [group.jac_identity() for _ in range((1 << c) - 1)]
        for base, scalar in pairs:
            if scalar == 1:
                if w_start == 0:
                    res = group.jac_add_mixed(res, base)
            else:
                d = (scalar >> w_start) % (1 << c)
                if d != 0:
                    buckets[d - 1] = group.jac_add_mixed(buckets[d - 1], base)
        running = group.jac_identity()
        for bkt in reversed(buckets):
            running = group.jac_add(running, bkt)
            res = group.jac_add(res, running)
        window_sums.append(res)
    lowest = window_sums[0]
    total = group.jac_identity()
    for ws in reversed(window_sums[1:]):
        total = group.jac_add(total, ws)
        for _ in range(c):
            total = group.jac_double(total)
    return group.jac_add(lowest, total)
