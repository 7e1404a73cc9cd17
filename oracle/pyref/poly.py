"""TEST INFRASTRUCTURE ONLY — CPU oracle: ark-poly 0.3 `Radix2EvaluationDomain` restated.

SURVEY.md Appendix C.4 (upstream ark-poly 0.3.0 `domain/radix2/*`, not vendored; reached from
`manta-crypto/src/arkworks/groth16.rs:597` through ark-groth16's `witness_map`, and directly
at `manta-trusted-setup/src/groth16/mpc.rs:367-381`).  Natural order in and out; `ifft`
includes the 1/m scaling; `coset_fft(x) = fft(x[i] * g^i)`; `coset_ifft(x) = ifft(x)[i] * g^-i`.
"""
from __future__ import annotations


class Radix2Domain:
    def __init__(self, curve, min_size: int):
        self.r = curve.r
        self.size = 1
        self.log_size = 0
        while self.size < min_size:
            self.size <<= 1
            self.log_size += 1
        assert self.log_size <= curve.two_adicity
        self.group_gen = pow(curve.root_of_unity, 1 << (curve.two_adicity - self.log_size), self.r)
        self.group_gen_inv = pow(self.group_gen, -1, self.r)
        self.size_inv = pow(self.size, -1, self.r)
        self.generator = curve.fr_gen
        self.generator_inv = pow(curve.fr_gen, -1, self.r)

    def _serial_fft(self, a, omega):
        """In-place Cooley-Tukey: bit-reverse then log n butterfly layers (ark `serial_radix2_fft`)."""
        r, n, log_n = self.r, self.size, self.log_size
        assert len(a) == n
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
                    t = a[k + j + m] * ws[j] % r
                    a[k + j + m] = (a[k + j] - t) % r
                    a[k + j] = (a[k + j] + t) % r
            m *= 2

    def _pad(self, a):
        assert len(a) <= self.size
        return list(a) + [0] * (self.size - len(a))

    def fft(self, a):
        a = self._pad(a)
        self._serial_fft(a, self.group_gen)
        return a

    def ifft(self, a):
        a = self._pad(a)
        self._serial_fft(a, self.group_gen_inv)
        return [x * self.size_inv % self.r for x in a]

    @staticmethod
    def _distribute_powers(a, g, r):
        p = 1
        out = []
        for x in a:
            out.append(x * p % r)
            p = p * g % r
        return out

    def coset_fft(self, a):
        return self.fft(self._distribute_powers(self._pad(a), self.generator, self.r))

    def coset_ifft(self, a):
        return self._distribute_powers(self.ifft(a), self.generator_inv, self.r)

    def vanishing_on_coset(self):
        """Z(g) = g^m - 1 (constant on the coset g*H)."""
        return (pow(self.generator, self.size, self.r) - 1) % self.r

    def lagrange_at(self, tau):
        """All L_j(tau), j < m (ark `evaluate_all_lagrange_coefficients`)."""
        r, m = self.r, self.size
        t_m = pow(tau, m, r)
        if t_m == 1:
            out, w = [0] * m, 1
            for j in range(m):
                if w == tau:
                    out[j] = 1
                    break
                w = w * self.group_gen % r
            return out
        # L_j(tau) = (tau^m - 1) * w^j / (m * (tau - w^j))
        z = (t_m - 1) * self.size_inv % r
        out, w = [], 1
        for _ in range(m):
            out.append(z * w % r * pow((tau - w) % r, -1, r) % r)
            w = w * self.group_gen % r
        return out
