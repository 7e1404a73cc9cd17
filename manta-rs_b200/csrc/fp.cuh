// Montgomery prime-field arithmetic on the sm_100a integer pipe (32-bit limbs, little-endian).
//
// Replaces, on the device, ark-ff 0.3 `Fp256`/`Fp384` (the arithmetic under
// manta-crypto/src/arkworks/ff.rs:24-75; upstream crate un-vendored).
// Field elements are kept in Montgomery form (value * R mod p, R = 2^(32 N)) and fully reduced
// (< p) between operations, so equality and the final canonical conversion are plain limb compares.
//
// Multiply: operand-scanning CIOS with the accumulator split into an "even-position" and an
// "odd-position" limb array, so that every 32x32->64 product is added by one carry-chained
// mad.lo.cc/madc.hi.cc pair — ptxas fuses each pair into a single IMAD.WIDE.U32(.X) with the carry
// in a predicate.  Cost per product: 2 N^2 wide MACs + N low multiplies (N = 12: 300; N = 8: 136).
#pragma once
#include <cstdint>
#include "bls12_381_constants.cuh"

#define MP_DEV __device__ __forceinline__
// out-of-line device code for everything off the hot loops (keeps cicc/ptxas time and code size sane)
#define MP_COLD __device__ __noinline__

namespace mp {

// ---- carry-chain primitives (PTX condition-code register; keep statements adjacent) ----------
MP_DEV void mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    asm volatile("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=&r"(lo), "=r"(hi) : "r"(a), "r"(b));
}
// (lo,hi) += a*b            -> CC
MP_DEV void mad_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
// (lo,hi) += a*b + CC       -> CC
MP_DEV void madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
// (lo,hi) = a*b + (clo,chi) + CC -> CC
MP_DEV void madc_wide_cc_from(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi) {
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %4; madc.hi.cc.u32 %1, %2, %3, %5;" : "=&r"(lo), "=r"(hi) : "r"(a), "r"(b), "r"(clo), "r"(chi));
}
// (lo,hi) = a*b + CC  (top pair of a shifted chain; cannot carry out)
MP_DEV void madc_wide_top(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    asm volatile("madc.lo.cc.u32 %0, %2, %3, 0; madc.hi.u32 %1, %2, %3, 0;" : "=&r"(lo), "=r"(hi) : "r"(a), "r"(b));
}
MP_DEV void add_cc(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
MP_DEV void addc_cc(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
MP_DEV void addc(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
MP_DEV void sub_cc(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
MP_DEV void subc_cc(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
MP_DEV void subc(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }

// ---- field parameter packs ---------------------------------------------------------------------
struct FqParams {
    static constexpr int N = 12;
    static constexpr uint32_t M0 = FQ_M0;
    static constexpr int BITS = 381;
    MP_DEV static const uint32_t* mod() { return FQ_MOD; }
    MP_DEV static const uint32_t* one() { return FQ_ONE; }
    MP_DEV static const uint32_t* r2() { return FQ_R2; }
    MP_DEV static const uint32_t* pm2() { return FQ_PM2; }
    MP_DEV static const uint32_t* half() { return FQ_HALF; }
    MP_DEV static const uint32_t* pshift() { return FQ_PSHIFT; }
    static constexpr int INV_ITERS = FQ_INV_ITERS;
    MP_DEV static const uint32_t* invfix() { return FQ_INVFIX; }
};
struct FrParams {
    static constexpr int N = 8;
    static constexpr uint32_t M0 = FR_M0;
    static constexpr int BITS = 255;
    MP_DEV static const uint32_t* mod() { return FR_MOD; }
    MP_DEV static const uint32_t* one() { return FR_ONE; }
    MP_DEV static const uint32_t* r2() { return FR_R2; }
    MP_DEV static const uint32_t* pm2() { return FR_PM2; }
    MP_DEV static const uint32_t* half() { return FR_HALF; }
    MP_DEV static const uint32_t* pshift() { return FR_PSHIFT; }
    static constexpr int INV_ITERS = FR_INV_ITERS;
    MP_DEV static const uint32_t* invfix() { return FR_INVFIX; }
};

template <class P>
struct Fp {
    static constexpr int N = P::N;
    uint32_t l[N];

    // ---- constructors -------------------------------------------------------------------------
    MP_DEV static Fp zero() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = 0;
        return r;
    }
    MP_DEV static Fp one() { return from_const(P::one()); }
    MP_DEV static Fp from_const(const uint32_t* c) {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = c[i];
        return r;
    }
    // 16-byte vector loads/stores (pointers must be 16-byte aligned)
    MP_DEV static Fp load(const void* p) {
        Fp r;
        const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
        for (int i = 0; i < N / 4; i++) {
            uint4 v = q[i];
            r.l[4 * i] = v.x; r.l[4 * i + 1] = v.y; r.l[4 * i + 2] = v.z; r.l[4 * i + 3] = v.w;
        }
        return r;
    }
    MP_DEV static Fp load_ro(const void* p) {  // read-only path
        Fp r;
        const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
        for (int i = 0; i < N / 4; i++) {
            uint4 v = __ldg(q + i);
            r.l[4 * i] = v.x; r.l[4 * i + 1] = v.y; r.l[4 * i + 2] = v.z; r.l[4 * i + 3] = v.w;
        }
        return r;
    }
    MP_DEV void store(void* p) const {
        uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
        for (int i = 0; i < N / 4; i++) q[i] = make_uint4(l[4 * i], l[4 * i + 1], l[4 * i + 2], l[4 * i + 3]);
    }

    // ---- predicates ---------------------------------------------------------------------------
    MP_DEV bool is_zero() const {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < N; i++) acc |= l[i];
        return acc == 0;
    }
    MP_DEV bool operator==(const Fp& o) const {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < N; i++) acc |= l[i] ^ o.l[i];
        return acc == 0;
    }
    MP_DEV bool operator!=(const Fp& o) const { return !(*this == o); }

    // ---- raw multi-limb helpers -----------------------------------------------------------------
    // r = a - b, returns borrow (1 if a < b)
    MP_DEV static uint32_t sub_raw(uint32_t* r, const uint32_t* a, const uint32_t* b) {
        sub_cc(r[0], a[0], b[0]);
#pragma unroll
        for (int i = 1; i < N; i++) subc_cc(r[i], a[i], b[i]);
        uint32_t borrow;
        subc(borrow, 0, 0);
        return borrow;  // 0 or 0xffffffff
    }
    // conditional final subtraction: x in [0, 2p) -> [0, p)
    MP_DEV void reduce_once() {
        uint32_t t[N];
        const uint32_t* m = P::mod();
        uint32_t borrow = sub_raw(t, l, m);
#pragma unroll
        for (int i = 0; i < N; i++) l[i] = borrow ? l[i] : t[i];
    }

    // ---- additive group -------------------------------------------------------------------------
    MP_DEV Fp operator+(const Fp& o) const {
        Fp r;
        add_cc(r.l[0], l[0], o.l[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) addc_cc(r.l[i], l[i], o.l[i]);
        addc(r.l[N - 1], l[N - 1], o.l[N - 1]);  // 2p < 2^(32N): no carry out
        r.reduce_once();
        return r;
    }
    MP_DEV Fp operator-(const Fp& o) const {
        Fp r;
        uint32_t borrow = sub_raw(r.l, l, o.l);
        const uint32_t* m = P::mod();
        uint32_t t[N];
        add_cc(t[0], r.l[0], m[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) addc_cc(t[i], r.l[i], m[i]);
        addc(t[N - 1], r.l[N - 1], m[N - 1]);
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = borrow ? t[i] : r.l[i];
        return r;
    }
    MP_DEV Fp neg() const {
        Fp r;
        const uint32_t* m = P::mod();
        sub_raw(r.l, m, l);
        bool z = is_zero();
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = z ? 0u : r.l[i];
        return r;
    }
    MP_DEV Fp dbl() const { return *this + *this; }
    MP_DEV static Fp select(bool c, const Fp& a, const Fp& b) {  // c ? a : b
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = c ? a.l[i] : b.l[i];
        return r;
    }

    // ---- Montgomery multiplication ----------------------------------------------------------------
    // One CIOS row on the split accumulator.  `x` lives at even limb positions (x[0] = position 0),
    // `y` at odd positions (y[k] = position k + 1).  Adds v * c where v has N limbs.
    //   y-chain first (cannot carry out), then x-chain whose carry lands in y[N-1].
    MP_DEV static void row_add(uint32_t* x, uint32_t* y, const uint32_t* v, uint32_t c) {
        mad_wide_cc(y[0], y[1], v[1], c);
#pragma unroll
        for (int j = 3; j < N; j += 2) madc_wide_cc(y[j - 1], y[j], v[j], c);
        // y chain never carries out of position N (bound argument in DESIGN.md); the CC it leaves is
        // overwritten by the next chain start.
        mad_wide_cc(x[0], x[1], v[0], c);
#pragma unroll
        for (int j = 2; j < N; j += 2) madc_wide_cc(x[j], x[j + 1], v[j], c);
        addc(y[N - 1], y[N - 1], 0);
    }

    MP_DEV Fp mul_cios(const Fp& o) const {
        const uint32_t* m = P::mod();
        // two role-swapping accumulators; after each row the value is divided by 2^32, which turns the
        // odd-position array into the even-position one and shifts the other by two limbs.
        uint32_t u[N], w[N];
        // ---- row 0: plain products ----
#pragma unroll
        for (int j = 0; j < N; j += 2) {
            mul_wide(u[j], u[j + 1], l[j], o.l[0]);      // even positions
            mul_wide(w[j], w[j + 1], l[j + 1], o.l[0]);  // odd positions
        }
        {
            uint32_t mi = u[0] * P::M0;
            row_add(u, w, m, mi);
        }
        // now u[0] == 0 and the pending shift makes w the even-position array.
#pragma unroll
        for (int i = 1; i < N; i++) {
            uint32_t* x = (i & 1) ? w : u;   // becomes even-position array
            uint32_t* yo = (i & 1) ? u : w;  // old even array: yo[0] == 0, yo[1] folds into x[0], rest shifts by 2
            uint32_t bi = o.l[i];
            // fold + shifted y-chain:  y'[k] = yo[k + 2]
            add_cc(x[0], x[0], yo[1]);
#pragma unroll
            for (int j = 1; j < N - 1; j += 2) madc_wide_cc_from(yo[j - 1], yo[j], l[j], bi, yo[j + 1], yo[j + 2]);
            madc_wide_top(yo[N - 2], yo[N - 1], l[N - 1], bi);
            // x-chain
            mad_wide_cc(x[0], x[1], l[0], bi);
#pragma unroll
            for (int j = 2; j < N; j += 2) madc_wide_cc(x[j], x[j + 1], l[j], bi);
            addc(yo[N - 1], yo[N - 1], 0);
            uint32_t mi = x[0] * P::M0;
            row_add(x, yo, m, mi);
        }
        // merge: result[k] = y'[k] + x[k + 1]
        uint32_t* x = ((N - 1) & 1) ? w : u;  // even-position array of the last row (x[0] == 0)
        uint32_t* y = ((N - 1) & 1) ? u : w;
        Fp r;
        add_cc(r.l[0], y[0], x[1]);
#pragma unroll
        for (int k = 1; k < N - 1; k++) addc_cc(r.l[k], y[k], x[k + 1]);
        addc(r.l[N - 1], y[N - 1], 0);
        r.reduce_once();
        return r;
    }

    // ---- wide (unreduced) products and Montgomery reduction ------------------------------------------
    // The production multiply is  redc(a (x) b)  with the 2N-limb product built by one level of (subtractive)
    // Karatsuba over N/2-limb halves: 3 (N/2)^2 + N^2 wide MACs instead of 2 N^2 (N = 12: 252 vs 288).  The
    // multiplier pipe (IMAD.WIDE: one warp instruction per 4 cycles per sub-partition) is the bottleneck of
    // every MSM kernel while the ALU pipe idles, so trading MACs for IADD3s is a net win.  Keeping the product
    // unreduced also lets sums of products share ONE reduction (Fq2, the Y3 coordinate of the point formulas).
    struct Wide {
        uint32_t l[2 * N];
    };

    // schoolbook H x H -> 2H limbs with the even/odd split accumulators (no reduction)
    template <int H>
    MP_DEV static void mul_half(uint32_t* t, const uint32_t* a, const uint32_t* b) {
        uint32_t E[2 * H + 1], O[2 * H];
#pragma unroll
        for (int k = 0; k < 2 * H + 1; k++) E[k] = 0;
#pragma unroll
        for (int k = 0; k < 2 * H; k++) O[k] = 0;
        // row 0: plain products
#pragma unroll
        for (int j = 0; j < H; j += 2) {
            mul_wide(E[j], E[j + 1], a[j], b[0]);
            mul_wide(O[j], O[j + 1], a[j + 1], b[0]);
        }
#pragma unroll
        for (int i = 1; i < H; i++) {
            const uint32_t bi = b[i];
            if (i & 1) {
                // a_even * b_i lands on odd positions: O[i + j - 1], O[i + j]
                mad_wide_cc(O[i - 1], O[i], a[0], bi);
#pragma unroll
                for (int j = 2; j < H; j += 2) madc_wide_cc(O[i + j - 1], O[i + j], a[j], bi);
                addc(O[i + H - 1], O[i + H - 1], 0);
                // a_odd * b_i lands on even positions: E[i + j], E[i + j + 1]
                mad_wide_cc(E[i + 1], E[i + 2], a[1], bi);
#pragma unroll
                for (int j = 3; j < H; j += 2) madc_wide_cc(E[i + j], E[i + j + 1], a[j], bi);
                addc(E[i + H + 1], E[i + H + 1], 0);
            } else {
                mad_wide_cc(E[i], E[i + 1], a[0], bi);
#pragma unroll
                for (int j = 2; j < H; j += 2) madc_wide_cc(E[i + j], E[i + j + 1], a[j], bi);
                addc(E[i + H], E[i + H], 0);
                mad_wide_cc(O[i], O[i + 1], a[1], bi);
#pragma unroll
                for (int j = 3; j < H; j += 2) madc_wide_cc(O[i + j - 1], O[i + j], a[j], bi);
                addc(O[i + H], O[i + H], 0);
            }
        }
        // t = E + (O << 32)
        t[0] = E[0];
        add_cc(t[1], E[1], O[0]);
#pragma unroll
        for (int k = 2; k < 2 * H - 1; k++) addc_cc(t[k], E[k], O[k - 1]);
        addc(t[2 * H - 1], E[2 * H - 1], O[2 * H - 2]);
    }

    // d = x - y over H limbs; returns an all-ones mask when x < y, and then d = y - x (absolute difference)
    template <int H>
    MP_DEV static uint32_t abs_diff(uint32_t* d, const uint32_t* x, const uint32_t* y) {
        sub_cc(d[0], x[0], y[0]);
#pragma unroll
        for (int k = 1; k < H; k++) subc_cc(d[k], x[k], y[k]);
        uint32_t mask;
        subc(mask, 0, 0);
        // conditional two's-complement negation: (d ^ mask) + (mask & 1)
        add_cc(d[0], d[0] ^ mask, mask & 1u);
#pragma unroll
        for (int k = 1; k < H - 1; k++) addc_cc(d[k], d[k] ^ mask, 0);
        addc(d[H - 1], d[H - 1] ^ mask, 0);
        return mask;
    }

    // t = a * b (2N limbs); a, b are plain N-limb integers (not necessarily reduced)
    MP_DEV static void mul_wide_full(uint32_t* t, const uint32_t* a, const uint32_t* b) {
        constexpr int H = N / 2;
        uint32_t z0[2 * H], z2[2 * H], zm[2 * H], da[H], db[H];
        mul_half<H>(z0, a, b);
        mul_half<H>(z2, a + H, b + H);
        uint32_t sa = abs_diff<H>(da, a, a + H);      // |a0 - a1|
        uint32_t sb = abs_diff<H>(db, b + H, b);      // |b1 - b0|
        mul_half<H>(zm, da, db);
        const uint32_t neg = sa ^ sb;                 // (a0 - a1)(b1 - b0) is negative
        // mid = z0 + z2 +- zm = a0 b1 + a1 b0   (2H + 1 limbs)
        uint32_t mid[2 * H + 1];
        add_cc(mid[0], z0[0], z2[0]);
#pragma unroll
        for (int k = 1; k < 2 * H; k++) addc_cc(mid[k], z0[k], z2[k]);
        addc(mid[2 * H], 0, 0);
        add_cc(mid[0], mid[0], neg & 1u);             // + 1 of the two's complement (cannot ripple past mid[2H])
#pragma unroll
        for (int k = 1; k < 2 * H; k++) addc_cc(mid[k], mid[k], 0);
        addc(mid[2 * H], mid[2 * H], 0);
        add_cc(mid[0], mid[0], zm[0] ^ neg);
#pragma unroll
        for (int k = 1; k < 2 * H; k++) addc_cc(mid[k], mid[k], zm[k] ^ neg);
        addc(mid[2 * H], mid[2 * H], neg);            // sign extension
        // t = z0 + (mid << 32H) + (z2 << 64H)
#pragma unroll
        for (int k = 0; k < H; k++) t[k] = z0[k];
        add_cc(t[H], z0[H], mid[0]);
#pragma unroll
        for (int k = 1; k < H; k++) addc_cc(t[H + k], z0[H + k], mid[k]);
#pragma unroll
        for (int k = 0; k < H; k++) addc_cc(t[2 * H + k], z2[k], mid[H + k]);
        addc_cc(t[3 * H], z2[H], mid[2 * H]);
#pragma unroll
        for (int k = 1; k < H - 1; k++) addc_cc(t[3 * H + k], z2[H + k], 0);
        addc(t[4 * H - 1], z2[2 * H - 1], 0);
    }

    MP_DEV static Wide mul_wide_w(const Fp& a, const Fp& b) {
        Wide w;
#ifdef MP_WIDE_KARATSUBA
        mul_wide_full(w.l, a.l, b.l);
#else
        mul_half<N>(w.l, a.l, b.l);  // schoolbook N x N
#endif
        return w;
    }
    // w = a + b on 2N limbs (caller guarantees no overflow: sums of a few products of reduced values)
    MP_DEV static Wide wide_add(const Wide& a, const Wide& b) {
        Wide w;
        add_cc(w.l[0], a.l[0], b.l[0]);
#pragma unroll
        for (int k = 1; k < 2 * N - 1; k++) addc_cc(w.l[k], a.l[k], b.l[k]);
        addc(w.l[2 * N - 1], a.l[2 * N - 1], b.l[2 * N - 1]);
        return w;
    }
    // w = a - b on 2N limbs (caller guarantees a >= b)
    MP_DEV static Wide wide_sub(const Wide& a, const Wide& b) {
        Wide w;
        sub_cc(w.l[0], a.l[0], b.l[0]);
#pragma unroll
        for (int k = 1; k < 2 * N - 1; k++) subc_cc(w.l[k], a.l[k], b.l[k]);
        subc(w.l[2 * N - 1], a.l[2 * N - 1], b.l[2 * N - 1]);
        return w;
    }
    // w = a - b + p * 2^bits(p): a multiple of p large enough (>= p^2) to keep a difference of two products of
    // reduced values non-negative, small enough (< R p / 8) not to disturb the bound of redc.
    MP_DEV static Wide wide_sub_biased(const Wide& a, const Wide& b) {
        Wide w;
        const uint32_t* ps = P::pshift();
        constexpr int Z = P::BITS / 32;  // low limbs of the bias that are zero
        sub_cc(w.l[0], a.l[0], b.l[0]);
#pragma unroll
        for (int k = 1; k < 2 * N - 1; k++) subc_cc(w.l[k], a.l[k], b.l[k]);
        subc(w.l[2 * N - 1], a.l[2 * N - 1], b.l[2 * N - 1]);
        add_cc(w.l[Z], w.l[Z], ps[Z]);
#pragma unroll
        for (int k = Z + 1; k < 2 * N - 1; k++) addc_cc(w.l[k], w.l[k], ps[k]);
        addc(w.l[2 * N - 1], w.l[2 * N - 1], ps[2 * N - 1]);
        return w;
    }
    // a + b without the conditional subtraction (< 2p, still fits N limbs); only as an operand of mul_wide_w
    MP_DEV static Fp add_noreduce(const Fp& a, const Fp& b) {
        Fp r;
        add_cc(r.l[0], a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) addc_cc(r.l[i], a.l[i], b.l[i]);
        addc(r.l[N - 1], a.l[N - 1], b.l[N - 1]);
        return r;
    }
    // generic "lazy" interface shared with Fq2 (used by the point formulas)
    MP_DEV static Wide mulw(const Fp& a, const Fp& b) { return mul_wide_w(a, b); }
    MP_DEV static Wide addw(const Wide& a, const Wide& b) { return wide_add(a, b); }
    MP_DEV static Fp redcw(const Wide& w) { return redc(w.l); }

    // Montgomery reduction: t / 2^(32 N) mod p for t < 2^(32 N) * p (result fully reduced).
    // Same split-accumulator rows as mul_cios with the a*b_i rows removed: N^2 wide MACs + N low multiplies.
    MP_DEV static Fp redc(const uint32_t* t) {
        const uint32_t* m = P::mod();
        uint32_t u[N], w[N];
#pragma unroll
        for (int j = 0; j < N; j++) u[j] = t[j];
        // ---- row 0: y array starts at zero, so its products need no carry chain
        {
            uint32_t mi = u[0] * P::M0;
#pragma unroll
            for (int j = 1; j < N; j += 2) mul_wide(w[j - 1], w[j], m[j], mi);
            mad_wide_cc(u[0], u[1], m[0], mi);
#pragma unroll
            for (int j = 2; j < N; j += 2) madc_wide_cc(u[j], u[j + 1], m[j], mi);
            addc(w[N - 1], w[N - 1], 0);
        }
#pragma unroll
        for (int i = 1; i < N; i++) {
            uint32_t* x = (i & 1) ? w : u;   // becomes the even-position array
            uint32_t* yo = (i & 1) ? u : w;  // old even array: yo[0] == 0, yo[1] folds into x[0], rest shifts by 2
            add_cc(x[0], x[0], yo[1]);
            uint32_t mi = x[0] * P::M0;
#pragma unroll
            for (int j = 1; j < N - 1; j += 2) madc_wide_cc_from(yo[j - 1], yo[j], m[j], mi, yo[j + 1], yo[j + 2]);
            madc_wide_top(yo[N - 2], yo[N - 1], m[N - 1], mi);
            mad_wide_cc(x[0], x[1], m[0], mi);
#pragma unroll
            for (int j = 2; j < N; j += 2) madc_wide_cc(x[j], x[j + 1], m[j], mi);
            addc(yo[N - 1], yo[N - 1], 0);
        }
        uint32_t* x = ((N - 1) & 1) ? w : u;
        uint32_t* y = ((N - 1) & 1) ? u : w;
        Fp r;
        add_cc(r.l[0], y[0], x[1]);
#pragma unroll
        for (int k = 1; k < N - 1; k++) addc_cc(r.l[k], y[k], x[k + 1]);
        addc(r.l[N - 1], y[N - 1], 0);
        // + high half of t
        add_cc(r.l[0], r.l[0], t[N]);
#pragma unroll
        for (int k = 1; k < N - 1; k++) addc_cc(r.l[k], r.l[k], t[N + k]);
        addc(r.l[N - 1], r.l[N - 1], t[2 * N - 1]);
        r.reduce_once();
        return r;
    }
    MP_DEV static Fp redc(const Wide& w) { return redc(w.l); }

#ifdef MP_MUL_KARATSUBA
    MP_DEV Fp operator*(const Fp& o) const {
        uint32_t t[2 * N];
        mul_wide_full(t, l, o.l);
        return redc(t);
    }
#else
    // measured on B200: the interleaved CIOS form wins — IADD3s are not free next to IMAD.WIDE (~1.2 clk each), so
    // Karatsuba's 36 saved MACs cost more than they save, and its 16 extra registers drop a resident block.
    MP_DEV Fp operator*(const Fp& o) const { return mul_cios(o); }
#endif
#ifdef MP_NO_FAST_SQR
    MP_DEV Fp sqr() const { return *this * *this; }
#else
    // Squaring: the N (N - 1) / 2 cross products once (same even/odd split accumulators as mul_half), doubled, plus the N
    // diagonal squares, then one Montgomery reduction: N (N + 1) / 2 + N^2 wide MACs instead of 2 N^2 (N = 12: 222 vs 288).
    MP_DEV static void sqr_wide(uint32_t* t, const uint32_t* a) {
        uint32_t E[2 * N + 2], O[2 * N + 2];  // E[p], E[p+1]: a product at even position p; O[p-1], O[p]: at odd position p
#pragma unroll
        for (int k = 0; k < 2 * N + 2; k++) { E[k] = 0; O[k] = 0; }
#pragma unroll
        for (int i = 0; i < N - 1; i++) {
            // j = i + 1, i + 3, ...: positions i + j of parity (2 i + 1) -> odd
            {
                const int p0 = 2 * i + 1;
                mad_wide_cc(O[p0 - 1], O[p0], a[i + 1], a[i]);
                int last = p0;
#pragma unroll
                for (int j = i + 3; j < N; j += 2) {
                    madc_wide_cc(O[i + j - 1], O[i + j], a[j], a[i]);
                    last = i + j;
                }
                addc(O[last + 1], O[last + 1], 0);
            }
            // j = i + 2, i + 4, ...: positions of parity (2 i) -> even
            if (i + 2 < N) {
                const int p0 = 2 * i + 2;
                mad_wide_cc(E[p0], E[p0 + 1], a[i + 2], a[i]);
                int last = p0;
#pragma unroll
                for (int j = i + 4; j < N; j += 2) {
                    madc_wide_cc(E[i + j], E[i + j + 1], a[j], a[i]);
                    last = i + j;
                }
                addc(E[last + 2], E[last + 2], 0);
            }
        }
        // cross = E + (O << 32), over 2N limbs
        uint32_t c[2 * N];
        c[0] = E[0];
        add_cc(c[1], E[1], O[0]);
#pragma unroll
        for (int k = 2; k < 2 * N - 1; k++) addc_cc(c[k], E[k], O[k - 1]);
        addc(c[2 * N - 1], E[2 * N - 1], O[2 * N - 2]);
        // t = 2 * cross + sum a_i^2 2^(64 i)
        uint32_t dlo, dhi;
        mul_wide(dlo, dhi, a[0], a[0]);
        t[0] = dlo;                                   // bit 0 of 2 * cross is 0 and c[0] == 0
        add_cc(t[1], dhi, c[1] << 1);
#pragma unroll
        for (int i = 1; i < N; i++) {
            mul_wide(dlo, dhi, a[i], a[i]);
            addc_cc(t[2 * i], dlo, __funnelshift_l(c[2 * i - 1], c[2 * i], 1));
            if (i < N - 1) addc_cc(t[2 * i + 1], dhi, __funnelshift_l(c[2 * i], c[2 * i + 1], 1));
            else addc(t[2 * i + 1], dhi, __funnelshift_l(c[2 * i], c[2 * i + 1], 1));
        }
    }
    MP_DEV Fp sqr() const {
        uint32_t t[2 * N];
        sqr_wide(t, l);
        return redc(t);
    }
#endif

    // ---- conversions ------------------------------------------------------------------------------
    MP_COLD Fp mul_cold(const Fp& o) const { return *this * o; }
    MP_DEV Fp to_mont() const { return *this * from_const(P::r2()); }
    MP_DEV Fp from_mont() const {
        Fp o = zero();
        o.l[0] = 1;
        return *this * o;
    }

    // ---- exponentiation / inversion (Fermat; off the hot loop) ----------------------------------------
    MP_COLD Fp pow_const(const uint32_t* e, int bits) const {
        Fp r = one();
        for (int i = bits - 1; i >= 0; i--) {
            r = r.sqr();
            if ((e[i >> 5] >> (i & 31)) & 1) r = r * *this;
        }
        return r;
    }
    MP_COLD Fp pow_u64(uint64_t e) const {
        Fp r = one();
        for (int i = 63; i >= 0; i--) {
            r = r.sqr();
            if ((e >> i) & 1) r = r * *this;
        }
        return r;
    }
    MP_DEV Fp inv() const { return pow_const(P::pm2(), P::BITS); }  // 0 -> 0

    // Inversion without multiplications: Kaliski's "almost Montgomery inverse" (binary extended Euclid whose cofactors
    // are only doubled, never halved mod p) runs on the ALU pipe and leaves the multiplier pipe - the MSM bottleneck - to
    // the other warps.  Phase 1 returns x = A^-1 2^k mod p for the limbs A of *this, BITS <= k <= 2 BITS; with A = a R the
    // Montgomery form of a^-1 is x R^2 2^-k = x 2^(2*32N - k), applied with two Montgomery products.  0 -> 0.
    // ---- word-level binary Euclid (the inversion of the batched-affine MSM) -------------------------------------
    // A single warp issues about one instruction per 4 cycles, so the latency of a serial algorithm is its instruction
    // count: the bit-level loop (inv_kaliski below, ~540 steps over 12-limb numbers) costs ~0.23 ms.  This variant follows
    // Pornin's "optimized binary GCD": 31 bit-steps at a time run on 64-bit approximations (low 31 bits + top 33 bits of
    // a, b) while recording the 2x2 update matrix (f0 g0; f1 g1), |entries| <= 2^31; the matrix is then applied once to the
    // full-length a, b (exact division by 2^31) and, mod p, to the cofactors u, v (one Montgomery word step, i.e. an extra
    // factor 2^-32).  Invariants: a 2^(31 i) = u 2^(32 i) A, b 2^(31 i) = v 2^(32 i) A (mod p).  After INV_ITERS =
    // ceil((2 bits - 1) / 31) rounds b = 1 and A^-1 = v 2^I; with A = a R the Montgomery form of a^-1 is
    // v 2^I R^2 = mont_mul(v, 2^I R^3).  No multiplier-pipe work except ~8 N wide MACs per round.  0 -> 0.

    // out = |a f + b g| / 2^31 (exact); returns true when a f + b g < 0
    MP_DEV static bool lin_comb_shift31(uint32_t (&out)[N], const uint32_t (&a)[N], const uint32_t (&b)[N], int64_t f, int64_t g) {
        const bool sf = f < 0, sg = g < 0;
        const uint32_t af = (uint32_t)(sf ? -f : f), ag = (uint32_t)(sg ? -g : g);
        uint32_t pp[N + 1], qq[N + 1], t[N + 1];
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < N; j++) { c += (uint64_t)a[j] * af; pp[j] = (uint32_t)c; c >>= 32; }
        pp[N] = (uint32_t)c;
        c = 0;
#pragma unroll
        for (int j = 0; j < N; j++) { c += (uint64_t)b[j] * ag; qq[j] = (uint32_t)c; c >>= 32; }
        qq[N] = (uint32_t)c;
        bool neg;
        if (sf == sg) {
            add_cc(t[0], pp[0], qq[0]);
#pragma unroll
            for (int j = 1; j < N; j++) addc_cc(t[j], pp[j], qq[j]);
            addc(t[N], pp[N], qq[N]);
            neg = sf;
        } else {
            sub_cc(t[0], pp[0], qq[0]);
#pragma unroll
            for (int j = 1; j <= N; j++) subc_cc(t[j], pp[j], qq[j]);
            uint32_t borrow;
            subc(borrow, 0, 0);
            if (borrow) {  // two's complement negation
                sub_cc(t[0], 0, t[0]);
#pragma unroll
                for (int j = 1; j < N; j++) subc_cc(t[j], 0, t[j]);
                subc(t[N], 0, t[N]);
            }
            neg = borrow ? !sf : sf;
        }
#pragma unroll
        for (int j = 0; j < N; j++) out[j] = __funnelshift_r(t[j], t[j + 1], 31);
        return neg;
    }
    // out = (u f + v g) / 2^32 mod p, in [0, p)
    MP_DEV static void mod_comb(uint32_t (&out)[N], const uint32_t (&u)[N], const uint32_t (&v)[N], int64_t f, int64_t g) {
        const uint32_t* m = P::mod();
        const bool sf = f < 0, sg = g < 0;
        const uint32_t af = (uint32_t)(sf ? -f : f), ag = (uint32_t)(sg ? -g : g);
        uint32_t uu[N], vv[N], t[N + 1];
        sub_raw(uu, m, u);  // -u = p - u (p itself when u = 0: still congruent, and the bounds below hold)
        sub_raw(vv, m, v);
#pragma unroll
        for (int j = 0; j < N; j++) { uu[j] = sf ? uu[j] : u[j]; vv[j] = sg ? vv[j] : v[j]; }
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < N; j++) {
            c += (uint64_t)uu[j] * af;
            uint64_t d = (uint64_t)vv[j] * ag;
            uint64_t lo = (c & 0xffffffffu) + (d & 0xffffffffu);
            t[j] = (uint32_t)lo;
            c = (c >> 32) + (d >> 32) + (lo >> 32);
        }
        t[N] = (uint32_t)c;  // uu af + vv ag < 2^(32N + 31)
        const uint32_t qm = t[0] * P::M0;
        c = 0;
#pragma unroll
        for (int j = 0; j < N; j++) {
            c += (uint64_t)m[j] * qm + t[j];
            t[j] = (uint32_t)c;
            c >>= 32;
        }
        c += t[N];
        // t[0] == 0 now; value / 2^32 = t[1..N-1], c  (< 2p)
        Fp r;
#pragma unroll
        for (int j = 0; j < N - 1; j++) r.l[j] = t[j + 1];
        r.l[N - 1] = (uint32_t)c;
        r.reduce_once();
#pragma unroll
        for (int j = 0; j < N; j++) out[j] = r.l[j];
    }
    MP_COLD Fp inv_gcd() const {
        if (is_zero()) return zero();
        const uint32_t* m = P::mod();
        uint32_t a[N], b[N], u[N], v[N];
#pragma unroll
        for (int i = 0; i < N; i++) { a[i] = l[i]; b[i] = m[i]; u[i] = 0; v[i] = 0; }
        u[0] = 1;
        for (int it = 0; it < P::INV_ITERS; it++) {
            // n = max(len(a), len(b), 64); 33 top bits from bit n - 33, 31 low bits
            uint32_t tw = 1, top = a[1] | b[1];
#pragma unroll
            for (int j = 2; j < N; j++) {
                const uint32_t w = a[j] | b[j];
                if (w) { tw = j; top = w; }
            }
            uint32_t n = 32 * tw + 32 - __clz(top);
            n = n < 64 ? 64 : n;
            const uint32_t s = n - 33, w0 = s >> 5, o = s & 31;
            uint32_t a0 = 0, a1 = 0, a2 = 0, b0 = 0, b1 = 0, b2 = 0;
#pragma unroll
            for (int j = 0; j < N; j++) {
                if ((uint32_t)j == w0) { a0 = a[j]; b0 = b[j]; }
                if ((uint32_t)j == w0 + 1) { a1 = a[j]; b1 = b[j]; }
                if ((uint32_t)j == w0 + 2) { a2 = a[j]; b2 = b[j]; }
            }
            uint64_t ah = (((uint64_t)a1 << 32) | a0) >> o, bh = (((uint64_t)b1 << 32) | b0) >> o;
            if (o) { ah |= (uint64_t)a2 << (64 - o); bh |= (uint64_t)b2 << (64 - o); }
            const uint64_t mask33 = (1ull << 33) - 1;
            uint64_t abar = (uint64_t)(a[0] & 0x7fffffffu) | ((ah & mask33) << 31);
            uint64_t bbar = (uint64_t)(b[0] & 0x7fffffffu) | ((bh & mask33) << 31);
            int64_t f0 = 1, g0 = 0, f1 = 0, g1 = 1;
#pragma unroll 1
            for (int j = 0; j < 31; j++) {
                const bool odd = abar & 1u;
                const bool sw = odd && (abar < bbar);
                const uint64_t ta = sw ? bbar : abar, tb = sw ? abar : bbar;
                const int64_t tf0 = sw ? f1 : f0, tf1 = sw ? f0 : f1, tg0 = sw ? g1 : g0, tg1 = sw ? g0 : g1;
                abar = (ta - (odd ? tb : 0)) >> 1;
                bbar = tb;
                f0 = tf0 - (odd ? tf1 : 0);
                g0 = tg0 - (odd ? tg1 : 0);
                f1 = tf1 << 1;
                g1 = tg1 << 1;
            }
            uint32_t na[N], nb[N], nu[N], nv[N];
            if (lin_comb_shift31(na, a, b, f0, g0)) { f0 = -f0; g0 = -g0; }
            if (lin_comb_shift31(nb, a, b, f1, g1)) { f1 = -f1; g1 = -g1; }
            mod_comb(nu, u, v, f0, g0);
            mod_comb(nv, u, v, f1, g1);
#pragma unroll
            for (int i = 0; i < N; i++) { a[i] = na[i]; b[i] = nb[i]; u[i] = nu[i]; v[i] = nv[i]; }
        }
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = v[i];
        return r.mul_cold(from_const(P::invfix()));
    }

    MP_DEV static void limbs_shr1(uint32_t (&a)[N]) {
#pragma unroll
        for (int i = 0; i < N - 1; i++) a[i] = __funnelshift_r(a[i], a[i + 1], 1);
        a[N - 1] >>= 1;
    }
    MP_DEV static void limbs_shl1(uint32_t (&a)[N]) {
#pragma unroll
        for (int i = N - 1; i > 0; i--) a[i] = __funnelshift_l(a[i - 1], a[i], 1);
        a[0] <<= 1;
    }
    MP_COLD Fp inv_kaliski() const {
        if (is_zero()) return zero();
        const uint32_t* m = P::mod();
        // u, v: the Euclid pair; r, s: cofactors (< 2p < 2^(32N)).  Everything stays in registers: no pointers, no
        // dynamic indices.  One of four cases per step; (u, r) and (v, s) swap roles between the symmetric ones.
        uint32_t u[N], v[N], r[N], s[N];
#pragma unroll
        for (int i = 0; i < N; i++) { u[i] = m[i]; v[i] = l[i]; r[i] = 0; s[i] = 0; }
        s[0] = 1;
        // (a branch-free formulation of the four cases was measured slower: every lane then pays for all of them)
        int k = 0;
        for (; k < 2 * 32 * N; k++) {
            uint32_t nz = 0;
#pragma unroll
            for (int i = 0; i < N; i++) nz |= v[i];
            if (!nz) break;
            if (!(u[0] & 1u)) {
                limbs_shr1(u);
                limbs_shl1(s);
            } else if (!(v[0] & 1u)) {
                limbs_shr1(v);
                limbs_shl1(r);
            } else {
                uint32_t t1[N], t2[N];
                uint32_t v_lt_u = sub_raw(t1, v, u);  // borrow: v < u
                sub_raw(t2, u, v);
                uint32_t sum[N];                       // r + s: the new value of whichever cofactor is not doubled
                add_cc(sum[0], r[0], s[0]);
#pragma unroll
                for (int i = 1; i < N - 1; i++) addc_cc(sum[i], r[i], s[i]);
                addc(sum[N - 1], r[N - 1], s[N - 1]);
                if (v_lt_u) {  // u > v: u = (u - v) / 2, r += s, s *= 2
#pragma unroll
                    for (int i = 0; i < N; i++) { u[i] = t2[i]; r[i] = sum[i]; }
                    limbs_shr1(u);
                    limbs_shl1(s);
                } else {       // v >= u: v = (v - u) / 2, s += r, r *= 2
#pragma unroll
                    for (int i = 0; i < N; i++) { v[i] = t1[i]; s[i] = sum[i]; }
                    limbs_shr1(v);
                    limbs_shl1(r);
                }
            }
        }
        // r < 2p: x = p - (r mod p)
        Fp x;
#pragma unroll
        for (int i = 0; i < N; i++) x.l[i] = r[i];
        x.reduce_once();
        x = x.neg();
        // x * 2^e with e = 2 * 32N - k: Montgomery products with R^2 (-> x R) and with the plain integer 2^e (-> x 2^e);
        // exponents that would not fit under p are taken out by doublings first
        int e = 2 * 32 * N - k;
        Fp y = x.mul_cold(from_const(P::r2()));
        while (e > P::BITS - 1) {
            y = y.dbl();
            e--;
        }
        Fp pw;
#pragma unroll
        for (int i = 0; i < N; i++) pw.l[i] = (i == (e >> 5)) ? (1u << (e & 31)) : 0u;
        return y.mul_cold(pw);
    }

    // canonical (non-Montgomery) integer > (p-1)/2 ?   (ark's `y > -y` flag)
    MP_DEV static bool canonical_gt_half(const Fp& canon) {
        uint32_t t[N];
        const uint32_t* h = P::half();
        uint32_t borrow = sub_raw(t, h, canon.l);  // half - canon < 0  <=> canon > half
        return borrow != 0;
    }
};

using Fq = Fp<FqParams>;
using Fr = Fp<FrParams>;

}  // namespace mp
