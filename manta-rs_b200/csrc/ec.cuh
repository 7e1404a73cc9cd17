// BLS12-381 group arithmetic on the device: Fq2 and short-Weierstrass points (a = 0) in
// affine and extended-Jacobian "XYZZ" coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2).
//
// Replaces ark-ec 0.3 `short_weierstrass_jacobian` (un-vendored upstream).  ark
// accumulates MSM buckets with the Jacobian mixed add (7M + 4S); XYZZ needs 8M + 2S and no
// per-step doubling of temporaries, and any coordinate system yields the same affine point, which is
// all that is observable in the proof bytes.
//
// Generic over the coordinate field F (Fq for G1, Fq2 for G2).
#pragma once
#include "fp.cuh"

// Lazy reduction (one Montgomery reduction for a sum of products) in the Fq2 multiply and in the Y3 coordinate of
// the addition formulas: +4.5 % on the G2 mixed add, neutral on G1 (tools/microbench/madd_bench.cu).  -DMP_NO_LAZY
// restores the plain forms.
#ifndef MP_NO_LAZY
#define MP_LAZY_FQ2 1
#define MP_LAZY_Y3 1
#endif

namespace mp {

// ---- Fq2 = Fq[u]/(u^2 + 1) -------------------------------------------------------------------------
struct Fq2 {
    Fq c0, c1;
    static constexpr int WORDS = 24;

    MP_DEV static Fq2 zero() { return {Fq::zero(), Fq::zero()}; }
    MP_DEV static Fq2 one() { return {Fq::one(), Fq::zero()}; }
    MP_DEV static Fq2 load(const void* p) {
        return {Fq::load(p), Fq::load(reinterpret_cast<const uint32_t*>(p) + 12)};
    }
    MP_DEV static Fq2 load_ro(const void* p) {
        return {Fq::load_ro(p), Fq::load_ro(reinterpret_cast<const uint32_t*>(p) + 12)};
    }
    MP_DEV void store(void* p) const {
        c0.store(p);
        c1.store(reinterpret_cast<uint32_t*>(p) + 12);
    }
    MP_DEV bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    MP_DEV bool operator==(const Fq2& o) const { return c0 == o.c0 && c1 == o.c1; }
    MP_DEV bool operator!=(const Fq2& o) const { return !(*this == o); }
    MP_DEV Fq2 operator+(const Fq2& o) const { return {c0 + o.c0, c1 + o.c1}; }
    MP_DEV Fq2 operator-(const Fq2& o) const { return {c0 - o.c0, c1 - o.c1}; }
    MP_DEV Fq2 neg() const { return {c0.neg(), c1.neg()}; }
    MP_DEV Fq2 dbl() const { return {c0.dbl(), c1.dbl()}; }
    // Unreduced product: d = a0 b0 - a1 b1 (+ bias), m = a0 b1 + a1 b0; three wide products (Karatsuba), and the
    // two Montgomery reductions are deferred so that sums of products can share them.
    struct Wide {
        Fq::Wide d, m;
    };
    MP_DEV static Wide mulw(const Fq2& a, const Fq2& b) {
        Fq::Wide t0 = Fq::mul_wide_w(a.c0, b.c0);
        Fq::Wide t1 = Fq::mul_wide_w(a.c1, b.c1);
        Fq::Wide t2 = Fq::mul_wide_w(Fq::add_noreduce(a.c0, a.c1), Fq::add_noreduce(b.c0, b.c1));
        Wide w;
        w.m = Fq::wide_sub(Fq::wide_sub(t2, t0), t1);
        w.d = Fq::wide_sub_biased(t0, t1);
        return w;
    }
    MP_DEV static Wide addw(const Wide& a, const Wide& b) { return {Fq::wide_add(a.d, b.d), Fq::wide_add(a.m, b.m)}; }
    MP_DEV static Fq2 redcw(const Wide& w) { return {Fq::redc(w.d), Fq::redc(w.m)}; }
#ifdef MP_LAZY_FQ2
    MP_DEV Fq2 operator*(const Fq2& o) const { return redcw(mulw(*this, o)); }
#else
    MP_DEV Fq2 operator*(const Fq2& o) const {  // Karatsuba: 3 Fq products
        Fq t0 = c0 * o.c0;
        Fq t1 = c1 * o.c1;
        Fq t2 = (c0 + c1) * (o.c0 + o.c1);
        return {t0 - t1, t2 - t0 - t1};
    }
#endif
    MP_DEV Fq2 sqr() const {  // (c0 + c1)(c0 - c1), 2 c0 c1
        Fq t = c0 * c1;
        return {(c0 + c1) * (c0 - c1), t.dbl()};
    }
    MP_COLD Fq2 inv() const {
        Fq n = (c0.sqr() + c1.sqr()).inv();
        return {c0 * n, (c1 * n).neg()};
    }
    MP_COLD Fq2 inv_gcd() const {  // same value as inv(), Fq inversion on the ALU pipe
        Fq n = (c0.mul_cold(c0) + c1.mul_cold(c1)).inv_gcd();
        return {c0.mul_cold(n), c1.mul_cold(n).neg()};
    }
    MP_DEV static Fq2 select(bool c, const Fq2& a, const Fq2& b) {
        return {Fq::select(c, a.c0, b.c0), Fq::select(c, a.c1, b.c1)};
    }
    MP_DEV Fq2 from_mont() const { return {c0.from_mont(), c1.from_mont()}; }
    MP_DEV Fq2 to_mont() const { return {c0.to_mont(), c1.to_mont()}; }
};

template <class F> struct FieldWords;
template <> struct FieldWords<Fq> { static constexpr int W = 12; };
template <> struct FieldWords<Fq2> { static constexpr int W = 24; };

// ---- points ------------------------------------------------------------------------------------------
// Affine point; (0, 0) encodes the point at infinity (never on y^2 = x^3 + b, b != 0).
template <class F>
struct Affine {
    F x, y;
    static constexpr int WORDS = 2 * FieldWords<F>::W;
    MP_DEV static Affine load(const void* p) {
        return {F::load(p), F::load(reinterpret_cast<const uint32_t*>(p) + FieldWords<F>::W)};
    }
    MP_DEV static Affine load_ro(const void* p) {
        return {F::load_ro(p), F::load_ro(reinterpret_cast<const uint32_t*>(p) + FieldWords<F>::W)};
    }
    MP_DEV void store(void* p) const {
        x.store(p);
        y.store(reinterpret_cast<uint32_t*>(p) + FieldWords<F>::W);
    }
    MP_DEV bool is_inf() const { return x.is_zero() && y.is_zero(); }
    MP_DEV static Affine inf() { return {F::zero(), F::zero()}; }
    MP_DEV Affine neg() const { return {x, y.neg()}; }
};

// Extended Jacobian; ZZ == 0 encodes infinity.
template <class F>
struct XYZZ {
    F X, Y, ZZ, ZZZ;
    static constexpr int WORDS = 4 * FieldWords<F>::W;

    MP_DEV static XYZZ inf() { return {F::zero(), F::zero(), F::zero(), F::zero()}; }
    MP_DEV bool is_inf() const { return ZZ.is_zero(); }
    MP_DEV static XYZZ from_affine(const Affine<F>& p) {
        if (p.is_inf()) return inf();
        return {p.x, p.y, F::one(), F::one()};
    }
    MP_DEV static XYZZ load(const void* p) {
        const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
        constexpr int W = FieldWords<F>::W;
        return {F::load(q), F::load(q + W), F::load(q + 2 * W), F::load(q + 3 * W)};
    }
    MP_DEV void store(void* p) const {
        uint32_t* q = reinterpret_cast<uint32_t*>(p);
        constexpr int W = FieldWords<F>::W;
        X.store(q); Y.store(q + W); ZZ.store(q + 2 * W); ZZZ.store(q + 3 * W);
    }
    MP_DEV XYZZ neg() const { return {X, Y.neg(), ZZ, ZZZ}; }

    // dbl-2008-s-1 (a = 0): 6M + 4S... counted as 2M + 5S + small ops below
    MP_COLD XYZZ dbl() const {
        if (is_inf()) return *this;
        F U = Y.dbl();
        F V = U.sqr();
        F W = U * V;
        F S = X * V;
        F M = X.sqr();
        M = M.dbl() + M;
        F X3 = M.sqr() - S.dbl();
        F Y3 = M * (S - X3) - W * Y;
        return {X3, Y3, V * ZZ, W * ZZZ};
    }
    // doubling of an affine point (mdbl-2008-s-1)
    MP_COLD static XYZZ dbl_affine(const Affine<F>& p) {
        if (p.is_inf()) return inf();
        F U = p.y.dbl();
        F V = U.sqr();
        F W = U * V;
        F S = p.x * V;
        F M = p.x.sqr();
        M = M.dbl() + M;
        F X3 = M.sqr() - S.dbl();
        F Y3 = M * (S - X3) - W * p.y;
        return {X3, Y3, V, W};
    }
    // madd-2008-s: 8M + 2S
    MP_DEV XYZZ add_mixed(const Affine<F>& p) const {
        if (p.is_inf()) return *this;
        if (is_inf()) return {p.x, p.y, F::one(), F::one()};
        F U2 = p.x * ZZ;
        F S2 = p.y * ZZZ;
        F P = U2 - X;
        F R = S2 - Y;
        if (P.is_zero()) {
            if (R.is_zero()) return dbl_affine(p);
            return inf();
        }
        F PP = P.sqr();
        F PPP = P * PP;
        F Q = X * PP;
        F X3 = R.sqr() - PPP - Q.dbl();
#ifdef MP_LAZY_Y3
        F Y3 = F::redcw(F::addw(F::mulw(R, Q - X3), F::mulw(Y.neg(), PPP)));  // one reduction for both products
#else
        F Y3 = R * (Q - X3) - Y * PPP;
#endif
        return {X3, Y3, ZZ * PP, ZZZ * PPP};
    }
    MP_COLD XYZZ add_mixed_cold(const Affine<F>& p) const { return add_mixed(p); }
    // add-2008-s: 12M + 2S
    MP_COLD XYZZ add(const XYZZ& o) const {
        if (o.is_inf()) return *this;
        if (is_inf()) return o;
        F U1 = X * o.ZZ;
        F U2 = o.X * ZZ;
        F S1 = Y * o.ZZZ;
        F S2 = o.Y * ZZZ;
        F P = U2 - U1;
        F R = S2 - S1;
        if (P.is_zero()) {
            if (R.is_zero()) return dbl();
            return inf();
        }
        F PP = P.sqr();
        F PPP = P * PP;
        F Q = U1 * PP;
        F X3 = R.sqr() - PPP - Q.dbl();
#ifdef MP_LAZY_Y3
        F Y3 = F::redcw(F::addw(F::mulw(R, Q - X3), F::mulw(S1.neg(), PPP)));
#else
        F Y3 = R * (Q - X3) - S1 * PPP;
#endif
        return {X3, Y3, ZZ * o.ZZ * PP, ZZZ * o.ZZZ * PPP};
    }
    // x = X/ZZ, y = Y/ZZZ with one inversion: i = (ZZ*ZZZ)^-1, 1/ZZ = i*ZZZ, 1/ZZZ = i*ZZ
    MP_COLD Affine<F> to_affine() const {
        if (is_inf()) return Affine<F>::inf();
        F i = (ZZ * ZZZ).inv_gcd();  // word-level binary Euclid: ~4x shorter than the Fermat ladder, off the multiplier pipe
        return {X * (i * ZZZ), Y * (i * ZZ)};
    }
};

// ---- 256-bit global accesses (sm_100: LDG / STG.E.ENL2.256) ---------------------------------------------------------
// Points are 96 / 192 bytes at 32-byte aligned addresses (window tables, point buffers), so a point moves in 3 / 6 of these
// instead of 6 / 12 LDG.128, and an x-coordinate in 2 / 3: half the load/store-unit requests and sector look-ups for the
// random gathers of the batched-affine kernels, which are bound there, not on bytes.  N words, N % 4 == 0; p 32-byte aligned.
template <int N, bool RO>
MP_DEV void load_words_256(uint32_t* r, const uint32_t* p) {
#pragma unroll
    for (int i = 0; i + 8 <= N; i += 8) {
        if (RO)
            asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(r[i]), "=r"(r[i + 1]), "=r"(r[i + 2]), "=r"(r[i + 3]), "=r"(r[i + 4]), "=r"(r[i + 5]), "=r"(r[i + 6]), "=r"(r[i + 7])
                         : "l"(p + i));
        else
            asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(r[i]), "=r"(r[i + 1]), "=r"(r[i + 2]), "=r"(r[i + 3]), "=r"(r[i + 4]), "=r"(r[i + 5]), "=r"(r[i + 6]), "=r"(r[i + 7])
                         : "l"(p + i)
                         : "memory");
    }
    if (N % 8) {
        const uint4 v = RO ? __ldg(reinterpret_cast<const uint4*>(p + N - 4)) : *reinterpret_cast<const uint4*>(p + N - 4);
        r[N - 4] = v.x; r[N - 3] = v.y; r[N - 2] = v.z; r[N - 1] = v.w;
    }
}
template <int N>
MP_DEV void store_words_256(uint32_t* p, const uint32_t* r) {
#pragma unroll
    for (int i = 0; i + 8 <= N; i += 8)
        asm volatile("st.global.v8.u32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};" ::"r"(r[i]), "r"(r[i + 1]), "r"(r[i + 2]), "r"(r[i + 3]), "r"(r[i + 4]),
                     "r"(r[i + 5]), "r"(r[i + 6]), "r"(r[i + 7]), "l"(p + i)
                     : "memory");
    if (N % 8) *reinterpret_cast<uint4*>(p + N - 4) = make_uint4(r[N - 4], r[N - 3], r[N - 2], r[N - 1]);
}
MP_DEV Fq field_from_words(const uint32_t* w, const Fq*) {
    Fq r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = w[i];
    return r;
}
MP_DEV Fq2 field_from_words(const uint32_t* w, const Fq2*) { return {field_from_words(w, (const Fq*)nullptr), field_from_words(w + 12, (const Fq*)nullptr)}; }
MP_DEV void field_to_words(uint32_t* w, const Fq& v) {
#pragma unroll
    for (int i = 0; i < 12; i++) w[i] = v.l[i];
}
MP_DEV void field_to_words(uint32_t* w, const Fq2& v) {
    field_to_words(w, v.c0);
    field_to_words(w + 12, v.c1);
}
// whole affine point / its x-coordinate / a field element, at a 32-byte aligned address
template <class F, bool RO>
MP_DEV Affine<F> load_point_256(const uint32_t* p) {
    constexpr int W = FieldWords<F>::W;
    uint32_t w[2 * W];
    load_words_256<2 * W, RO>(w, p);
    return {field_from_words(w, (const F*)nullptr), field_from_words(w + W, (const F*)nullptr)};
}
template <class F, bool RO>
MP_DEV F load_field_256(const uint32_t* p) {
    constexpr int W = FieldWords<F>::W;
    uint32_t w[W];
    load_words_256<W, RO>(w, p);
    return field_from_words(w, (const F*)nullptr);
}
template <class F>
MP_DEV void store_point_256(uint32_t* p, const Affine<F>& a) {
    constexpr int W = FieldWords<F>::W;
    uint32_t w[2 * W];
    field_to_words(w, a.x);
    field_to_words(w + W, a.y);
    store_words_256<2 * W>(p, w);
}

using G1Affine = Affine<Fq>;
using G2Affine = Affine<Fq2>;
using G1XYZZ = XYZZ<Fq>;
using G2XYZZ = XYZZ<Fq2>;

}  // namespace mp
