// TEST INFRASTRUCTURE ONLY — CPU oracle: C++ restatement of the reference's Groth16 proving path.
//
// The reference (`manta-crypto/src/arkworks/groth16.rs:588-600`) delegates to arkworks 0.3.0 crates that are not
// vendored under /root/reference (ark-groth16 / ark-ec / ark-ff / ark-poly / ark-serialize, pinned by
// `manta-crypto/Cargo.toml:76-87`), and no Rust toolchain exists in this image, so the algorithms are restated
// here from their published form:
//   * Fp256 / Fp384 Montgomery arithmetic on 64-bit limbs                         (ark-ff)
//   * Jacobian add-2007-bl / madd-2007-bl / dbl-2009-l                             (ark-ec short_weierstrass_jacobian)
//   * VariableBaseMSM::multi_scalar_mul: c = 3 if n < 32 else ceil(log2 n)*69/100+2, zero scalars dropped,
//     unit scalars added in window 0, 2^c - 1 buckets per window, running sum, Horner  (ark-ec msm/variable_base.rs)
//   * serial radix-2 FFT, coset (g = 7) transforms, witness_map                    (ark-poly, ark-groth16 r1cs_to_qap.rs)
//   * create_proof and the compressed proof encoding                              (ark-groth16 prover.rs, ark-serialize)
// Parity status: "parity unpinned" for proof BYTES (the reference holds no known-answer proof; every prove test
// draws from OsRng — SURVEY.md §8c).  Pinned pieces: Fr arithmetic by the reference's Poseidon fixtures, the
// point encoding by the reference's BN254 verifying-key files (tests/test_oracle_kats.py, via the Python twin
// oracle/pyref which this file is cross-checked against), and every proof by the trapdoor closed form.
//
// Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load this library.
// `threads` > 1 mirrors arkworks' optional `parallel` feature (windows / transforms in parallel) for the
// "all host cores" baseline; the reference as shipped is single-threaded (SURVEY.md §0 finding 5).
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;

// ------------------------------------------------------------------------------------------------------------
// prime fields
// ------------------------------------------------------------------------------------------------------------
template <int N>
struct FieldParams {
    uint64_t mod[N];
    uint64_t r1[N];   // R mod p
    uint64_t r2[N];   // R^2 mod p
    uint64_t inv;     // -p^-1 mod 2^64
    uint64_t half[N]; // (p-1)/2
    int bits;
};

template <int N>
static bool geq(const uint64_t* a, const uint64_t* b) {
    for (int i = N - 1; i >= 0; i--) {
        if (a[i] != b[i]) return a[i] > b[i];
    }
    return true;
}
template <int N>
static uint64_t add_n(uint64_t* r, const uint64_t* a, const uint64_t* b) {
    u128 c = 0;
    for (int i = 0; i < N; i++) {
        c += (u128)a[i] + b[i];
        r[i] = (uint64_t)c;
        c >>= 64;
    }
    return (uint64_t)c;
}
template <int N>
static uint64_t sub_n(uint64_t* r, const uint64_t* a, const uint64_t* b) {
    uint64_t borrow = 0;
    for (int i = 0; i < N; i++) {
        u128 d = (u128)a[i] - b[i] - borrow;
        r[i] = (uint64_t)d;
        borrow = (uint64_t)(d >> 64) & 1;
    }
    return borrow;
}

template <int N, const FieldParams<N>* (*PP)()>
struct Fp {
    uint64_t l[N];
    static const FieldParams<N>& P() { return *PP(); }
    static Fp zero() { Fp r; memset(r.l, 0, sizeof(r.l)); return r; }
    static Fp one() { Fp r; memcpy(r.l, P().r1, sizeof(r.l)); return r; }
    bool is_zero() const { for (int i = 0; i < N; i++) if (l[i]) return false; return true; }
    bool operator==(const Fp& o) const { return memcmp(l, o.l, sizeof(l)) == 0; }
    bool operator!=(const Fp& o) const { return !(*this == o); }
    Fp operator+(const Fp& o) const {
        Fp r;
        uint64_t c = add_n<N>(r.l, l, o.l);
        if (c || geq<N>(r.l, P().mod)) sub_n<N>(r.l, r.l, P().mod);
        return r;
    }
    Fp operator-(const Fp& o) const {
        Fp r;
        if (sub_n<N>(r.l, l, o.l)) add_n<N>(r.l, r.l, P().mod);
        return r;
    }
    Fp neg() const {
        if (is_zero()) return *this;
        Fp r;
        sub_n<N>(r.l, P().mod, l);
        return r;
    }
    Fp dbl() const { return *this + *this; }
    // CIOS Montgomery product with the "no-carry" merge of the two inner loops (valid because the top bit of
    // both moduli is clear) — the same loop structure ark-ff 0.3 generates for Fp256/Fp384.
    Fp operator*(const Fp& o) const {
        const FieldParams<N>& p = P();
        uint64_t r[N];
        for (int j = 0; j < N; j++) r[j] = 0;
#pragma GCC unroll 8
        for (int i = 0; i < N; i++) {
            u128 t = (u128)l[0] * o.l[i] + r[0];
            uint64_t carry1 = (uint64_t)(t >> 64);
            uint64_t k = (uint64_t)t * p.inv;
            u128 t2 = (u128)k * p.mod[0] + (uint64_t)t;
            uint64_t carry2 = (uint64_t)(t2 >> 64);
#pragma GCC unroll 8
            for (int j = 1; j < N; j++) {
                t = (u128)l[j] * o.l[i] + r[j] + carry1;
                carry1 = (uint64_t)(t >> 64);
                t2 = (u128)k * p.mod[j] + (uint64_t)t + carry2;
                carry2 = (uint64_t)(t2 >> 64);
                r[j - 1] = (uint64_t)t2;
            }
            r[N - 1] = carry1 + carry2;
        }
        Fp res;
        memcpy(res.l, r, sizeof(r));
        if (geq<N>(res.l, p.mod)) sub_n<N>(res.l, res.l, p.mod);
        return res;
    }
    Fp sqr() const { return *this * *this; }
    Fp pow(const uint64_t* e, int nlimbs) const {
        Fp r = one();
        for (int i = nlimbs * 64 - 1; i >= 0; i--) {
            r = r.sqr();
            if ((e[i / 64] >> (i % 64)) & 1) r = r * *this;
        }
        return r;
    }
    Fp inv() const {  // Fermat
        uint64_t e[N];
        uint64_t two[N] = {2};
        sub_n<N>(e, P().mod, two);
        return pow(e, N);
    }
    static Fp from_canonical(const uint64_t* c) {
        Fp r, r2;
        memcpy(r.l, c, sizeof(r.l));
        memcpy(r2.l, P().r2, sizeof(r2.l));
        return r * r2;
    }
    static Fp from_u64(uint64_t v) {
        uint64_t c[N] = {v};
        return from_canonical(c);
    }
    void to_canonical(uint64_t* c) const {
        Fp o = zero();
        o.l[0] = 1;
        Fp r = *this * o;
        memcpy(c, r.l, sizeof(r.l));
    }
    // ark `y > -y` on canonical integers
    bool lex_larger() const {
        uint64_t c[N];
        to_canonical(c);
        for (int i = N - 1; i >= 0; i--)
            if (c[i] != P().half[i]) return c[i] > P().half[i];
        return false;
    }
};

template <int N>
static void init_params(FieldParams<N>& p, const uint64_t* mod, int bits) {
    memcpy(p.mod, mod, sizeof(p.mod));
    p.bits = bits;
    uint64_t inv = 1;
    for (int i = 0; i < 63; i++) { inv *= inv; inv *= mod[0]; }  // mod[0]^(2^63 - 1) = mod[0]^-1 mod 2^64
    p.inv = (uint64_t)(0 - inv);
    // R mod p by doubling 1, 64*N times; R^2 by doubling another 64*N times
    uint64_t x[N] = {1};
    for (int i = 0; i < 2 * 64 * N; i++) {
        uint64_t c = add_n<N>(x, x, x);
        if (c || geq<N>(x, mod)) sub_n<N>(x, x, mod);
        if (i == 64 * N - 1) memcpy(p.r1, x, sizeof(x));
    }
    memcpy(p.r2, x, sizeof(x));
    uint64_t onev[N] = {1};
    sub_n<N>(p.half, mod, onev);
    for (int i = 0; i < N; i++) p.half[i] = (p.half[i] >> 1) | (i + 1 < N ? p.half[i + 1] << 63 : 0);
}

static const uint64_t FQ_MOD[6] = {0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
                                   0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL};
static const uint64_t FR_MOD[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL};

static FieldParams<6> g_fq;
static FieldParams<4> g_fr;
static const FieldParams<6>* fq_params() { return &g_fq; }
static const FieldParams<4>* fr_params() { return &g_fr; }
typedef Fp<6, fq_params> Fq;
typedef Fp<4, fr_params> Fr;

struct Fq2 {
    Fq c0, c1;
    static Fq2 zero() { return {Fq::zero(), Fq::zero()}; }
    static Fq2 one() { return {Fq::one(), Fq::zero()}; }
    bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    bool operator==(const Fq2& o) const { return c0 == o.c0 && c1 == o.c1; }
    bool operator!=(const Fq2& o) const { return !(*this == o); }
    Fq2 operator+(const Fq2& o) const { return {c0 + o.c0, c1 + o.c1}; }
    Fq2 operator-(const Fq2& o) const { return {c0 - o.c0, c1 - o.c1}; }
    Fq2 neg() const { return {c0.neg(), c1.neg()}; }
    Fq2 dbl() const { return {c0.dbl(), c1.dbl()}; }
    Fq2 operator*(const Fq2& o) const {
        Fq t0 = c0 * o.c0, t1 = c1 * o.c1;
        return {t0 - t1, (c0 + c1) * (o.c0 + o.c1) - t0 - t1};
    }
    Fq2 sqr() const { return {(c0 + c1) * (c0 - c1), (c0 * c1).dbl()}; }
    Fq2 inv() const {
        Fq n = (c0.sqr() + c1.sqr()).inv();
        return {c0 * n, (c1 * n).neg()};
    }
    bool lex_larger() const {  // Ord on Fq2: c1 first, then c0
        if (!c1.is_zero()) return c1.lex_larger();
        return c0.lex_larger();
    }
};

// ------------------------------------------------------------------------------------------------------------
// short Weierstrass (a = 0), Jacobian coordinates — formulas of ark-ec 0.3 short_weierstrass_jacobian.rs
// ------------------------------------------------------------------------------------------------------------
template <class F>
struct AffineT {
    F x, y;
    bool inf;
};
template <class F>
struct Jac {
    F X, Y, Z;
    static Jac identity() { return {F::zero(), F::one(), F::zero()}; }
    bool is_identity() const { return Z.is_zero(); }
    void dbl_in_place() {  // dbl-2009-l
        if (is_identity()) return;
        F A = X.sqr(), B = Y.sqr(), C = B.sqr();
        F D = ((X + B).sqr() - A - C).dbl();
        F E = A + A.dbl();
        F Fv = E.sqr();
        Z = (Z * Y).dbl();
        X = Fv - D - D;
        Y = (D - X) * E - C.dbl().dbl().dbl();
    }
    void add_mixed(const AffineT<F>& o) {  // madd-2007-bl
        if (o.inf) return;
        if (is_identity()) { X = o.x; Y = o.y; Z = F::one(); return; }
        F Z1Z1 = Z.sqr();
        F U2 = o.x * Z1Z1;
        F S2 = (o.y * Z) * Z1Z1;
        if (X == U2 && Y == S2) { dbl_in_place(); return; }
        F H = U2 - X;
        F HH = H.sqr();
        F I = HH.dbl().dbl();
        F J = H * I;
        F r = (S2 - Y).dbl();
        F V = X * I;
        F X3 = r.sqr() - J - V.dbl();
        F Y3 = r * (V - X3) - (Y * J).dbl();
        F Z3 = (Z + H).sqr() - Z1Z1 - HH;
        X = X3; Y = Y3; Z = Z3;
    }
    void add(const Jac& o) {  // add-2007-bl
        if (is_identity()) { *this = o; return; }
        if (o.is_identity()) return;
        F Z1Z1 = Z.sqr(), Z2Z2 = o.Z.sqr();
        F U1 = X * Z2Z2, U2 = o.X * Z1Z1;
        F S1 = Y * o.Z * Z2Z2, S2 = o.Y * Z * Z1Z1;
        if (U1 == U2 && S1 == S2) { dbl_in_place(); return; }
        F H = U2 - U1;
        F I = H.dbl().sqr();
        F J = H * I;
        F r = (S2 - S1).dbl();
        F V = U1 * I;
        F X3 = r.sqr() - J - V.dbl();
        F Y3 = r * (V - X3) - (S1 * J).dbl();
        F Z3 = ((Z + o.Z).sqr() - Z1Z1 - Z2Z2) * H;
        X = X3; Y = Y3; Z = Z3;
    }
    Jac neg() const { return {X, Y.neg(), Z}; }
    AffineT<F> to_affine() const {
        if (is_identity()) return {F::zero(), F::zero(), true};
        F zi = Z.inv();
        F zi2 = zi.sqr();
        return {X * zi2, Y * zi2 * zi, false};
    }
    // double-and-add, MSB first (ark `mul` on a BigInteger)
    Jac mul(const uint64_t* k, int nlimbs) const {
        Jac r = identity();
        bool started = false;
        for (int i = nlimbs * 64 - 1; i >= 0; i--) {
            if (started) r.dbl_in_place();
            if ((k[i / 64] >> (i % 64)) & 1) { r.add(*this); started = true; }
        }
        return r;
    }
};
typedef AffineT<Fq> G1A;
typedef AffineT<Fq2> G2A;
typedef Jac<Fq> G1J;
typedef Jac<Fq2> G2J;

template <class F>
static Jac<F> from_affine(const AffineT<F>& a) {
    if (a.inf) return Jac<F>::identity();
    return {a.x, a.y, F::one()};
}

// ------------------------------------------------------------------------------------------------------------
// ark-serialize 0.3
// ------------------------------------------------------------------------------------------------------------
static Fq read_fq(const uint8_t* p, bool strip) {
    uint64_t c[6];
    memcpy(c, p, 48);
    if (strip) c[5] &= 0x3fffffffffffffffULL;
    return Fq::from_canonical(c);
}
static void write_fq(uint8_t* p, const Fq& v) {
    uint64_t c[6];
    v.to_canonical(c);
    memcpy(p, c, 48);
}
static G1A read_g1_uncompressed(const uint8_t* p) {
    if (p[95] & 0x40) return {Fq::zero(), Fq::zero(), true};
    return {read_fq(p, false), read_fq(p + 48, true), false};
}
static G2A read_g2_uncompressed(const uint8_t* p) {
    if (p[191] & 0x40) return {Fq2::zero(), Fq2::zero(), true};
    return {{read_fq(p, false), read_fq(p + 48, false)}, {read_fq(p + 96, false), read_fq(p + 144, true)}, false};
}
static void write_g1_uncompressed(uint8_t* p, const G1A& a) {
    memset(p, 0, 96);
    if (a.inf) { p[95] = 0x40; return; }
    write_fq(p, a.x);
    write_fq(p + 48, a.y);
}
static void write_g2_uncompressed(uint8_t* p, const G2A& a) {
    memset(p, 0, 192);
    if (a.inf) { p[191] = 0x40; return; }
    write_fq(p, a.x.c0); write_fq(p + 48, a.x.c1); write_fq(p + 96, a.y.c0); write_fq(p + 144, a.y.c1);
}
static void write_g1_compressed(uint8_t* p, const G1A& a) {
    memset(p, 0, 48);
    if (a.inf) { p[47] = 0x40; return; }
    write_fq(p, a.x);
    if (a.y.lex_larger()) p[47] |= 0x80;
}
static void write_g2_compressed(uint8_t* p, const G2A& a) {
    memset(p, 0, 96);
    if (a.inf) { p[95] = 0x40; return; }
    write_fq(p, a.x.c0);
    write_fq(p + 48, a.x.c1);
    if (a.y.lex_larger()) p[95] |= 0x80;
}

// ------------------------------------------------------------------------------------------------------------
// VariableBaseMSM::multi_scalar_mul (ark-ec 0.3) — scalars are canonical 4-limb integers
// ------------------------------------------------------------------------------------------------------------
static int ark_log2(size_t x) {  // ark_std::log2 = ceil(log2(x)), 0 for x <= 1
    if (x <= 1) return 0;
    int n = 0;
    size_t v = x - 1;
    while (v) { n++; v >>= 1; }
    return n;
}
static int ark_window(size_t size) { return size < 32 ? 3 : ark_log2(size) * 69 / 100 + 2; }

static bool scalar_is_zero(const uint64_t* s) { return !(s[0] | s[1] | s[2] | s[3]); }
static bool scalar_is_one(const uint64_t* s) { return s[0] == 1 && !(s[1] | s[2] | s[3]); }
static uint64_t scalar_window(const uint64_t* s, int start, int c) {  // (s >> start) % 2^c
    int w = start / 64, o = start % 64;
    u128 v = s[w];
    if (w + 1 < 4) v |= (u128)s[w + 1] << 64;
    return (uint64_t)(v >> o) & (((uint64_t)1 << c) - 1);
}

// One window of ark's `multi_scalar_mul`: buckets of the c-bit digit starting at bit w_start, then the running-sum fold.
template <class F>
static Jac<F> msm_window(const AffineT<F>* bases, const uint64_t* scalars, size_t size, int c, int w_start) {
    Jac<F> res = Jac<F>::identity();
    std::vector<Jac<F>> buckets(((size_t)1 << c) - 1, Jac<F>::identity());
    for (size_t i = 0; i < size; i++) {
        const uint64_t* s = scalars + 4 * i;
        if (scalar_is_zero(s)) continue;
        if (scalar_is_one(s)) {
            if (w_start == 0) res.add_mixed(bases[i]);
        } else {
            uint64_t d = scalar_window(s, w_start, c);
            if (d) buckets[d - 1].add_mixed(bases[i]);
        }
    }
    Jac<F> running = Jac<F>::identity();
    for (size_t b = buckets.size(); b-- > 0;) {
        running.add(buckets[b]);
        res.add(running);
    }
    return res;
}
// ... and the combination of the window sums: lowest + sum_{w >= 1} 2^(c w) window_sums[w] by Horner
template <class F>
static Jac<F> msm_combine(const std::vector<Jac<F>>& window_sums, int c) {
    Jac<F> lowest = window_sums[0];
    Jac<F> total = Jac<F>::identity();
    for (size_t wi = window_sums.size() - 1; wi >= 1; wi--) {
        total.add(window_sums[wi]);
        for (int k = 0; k < c; k++) total.dbl_in_place();
    }
    lowest.add(total);
    return lowest;
}
static int msm_num_windows(int c) { return (255 + c - 1) / c; }

template <class F>
static Jac<F> msm(const AffineT<F>* bases, const uint64_t* scalars, size_t n_bases, size_t n_scalars, int threads) {
    size_t size = std::min(n_bases, n_scalars);
    const int c = ark_window(size);
    std::vector<Jac<F>> window_sums(msm_num_windows(c));
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) if (threads > 1)
    for (size_t wi = 0; wi < window_sums.size(); wi++) window_sums[wi] = msm_window<F>(bases, scalars, size, c, (int)wi * c);
    return msm_combine<F>(window_sums, c);
}

// ------------------------------------------------------------------------------------------------------------
// Radix2EvaluationDomain (ark-poly 0.3): serial in-place FFT, natural order in and out
// ------------------------------------------------------------------------------------------------------------
struct Domain {
    size_t size;
    int log_size;
    Fr group_gen, group_gen_inv, size_inv, generator, generator_inv;
};
static Fr fr_pow_u64(Fr b, uint64_t e) {
    uint64_t ee[1] = {e};
    return b.pow(ee, 1);
}
static Domain make_domain(size_t min_size) {
    Domain d;
    d.size = 1;
    d.log_size = 0;
    while (d.size < min_size) { d.size <<= 1; d.log_size++; }
    Fr g = Fr::from_u64(7);
    // 2^32-th root of unity: 7^((r-1)/2^32)
    uint64_t e[4];
    uint64_t onev[4] = {1};
    sub_n<4>(e, FR_MOD, onev);
    for (int i = 0; i < 4; i++) e[i] = (e[i] >> 32) | (i + 1 < 4 ? e[i + 1] << 32 : 0);
    Fr root = g.pow(e, 4);
    for (int i = d.log_size; i < 32; i++) root = root.sqr();
    d.group_gen = root;
    d.group_gen_inv = root.inv();
    d.size_inv = Fr::from_u64(d.size).inv();
    d.generator = g;
    d.generator_inv = g.inv();
    return d;
}
static void serial_fft(Fr* a, const Domain& d, const Fr& omega) {
    const size_t n = d.size;
    const int log_n = d.log_size;
    for (size_t k = 0; k < n; k++) {
        size_t rk = 0;
        for (int b = 0; b < log_n; b++) rk |= ((k >> b) & 1) << (log_n - 1 - b);
        if (k < rk) std::swap(a[k], a[rk]);
    }
    size_t m = 1;
    for (int s = 0; s < log_n; s++) {
        Fr w_m = fr_pow_u64(omega, n / (2 * m));
        for (size_t k = 0; k < n; k += 2 * m) {
            Fr w = Fr::one();
            for (size_t j = 0; j < m; j++) {
                Fr t = a[k + j + m] * w;
                a[k + j + m] = a[k + j] - t;
                a[k + j] = a[k + j] + t;
                w = w * w_m;
            }
        }
        m *= 2;
    }
}
static void distribute_powers(Fr* a, size_t n, const Fr& g) {
    Fr p = Fr::one();
    for (size_t i = 0; i < n; i++) { a[i] = a[i] * p; p = p * g; }
}
static void fft(Fr* a, const Domain& d) { serial_fft(a, d, d.group_gen); }
static void ifft(Fr* a, const Domain& d) {
    serial_fft(a, d, d.group_gen_inv);
    for (size_t i = 0; i < d.size; i++) a[i] = a[i] * d.size_inv;
}
static void coset_fft(Fr* a, const Domain& d) { distribute_powers(a, d.size, d.generator); fft(a, d); }
static void coset_ifft(Fr* a, const Domain& d) { ifft(a, d); distribute_powers(a, d.size, d.generator_inv); }

// ------------------------------------------------------------------------------------------------------------
// proving context: parsed key + matrices (Montgomery form)
// ------------------------------------------------------------------------------------------------------------
struct Matrix {
    std::vector<uint64_t> row_ptr;
    std::vector<uint32_t> col;
    std::vector<Fr> coeff;
};
struct OracleCtx {
    G1A alpha_g1, beta_g1, delta_g1;
    G2A beta_g2, delta_g2;
    std::vector<G1A> a_query, b_g1_query, h_query, l_query;
    std::vector<G2A> b_g2_query;
    uint64_t p = 0, w = 0, K = 0;
    Matrix mat[3];
    Domain dom;
};

static Fr eval_row(const Matrix& m, size_t row, const Fr* z) {
    Fr acc = Fr::zero();
    for (uint64_t e = m.row_ptr[row]; e < m.row_ptr[row + 1]; e++) acc = acc + m.coeff[e] * z[m.col[e]];
    return acc;
}

// R1CStoQAP::witness_map (ark-groth16 0.3), z in Montgomery form; returns h (m coefficients, Montgomery)
static std::vector<Fr> witness_map(const OracleCtx& c, const Fr* z, int threads) {
    const size_t m = c.dom.size, K = c.K, p = c.p;
    std::vector<Fr> a(m, Fr::zero()), b(m, Fr::zero()), cc(m, Fr::zero());
    for (size_t i = 0; i < K; i++) { a[i] = eval_row(c.mat[0], i, z); b[i] = eval_row(c.mat[1], i, z); }
    for (size_t j = 0; j < p; j++) a[K + j] = z[j];
    for (size_t i = 0; i < K; i++) cc[i] = eval_row(c.mat[2], i, z);
    Fr* vecs[3] = {a.data(), b.data(), cc.data()};
#pragma omp parallel for num_threads(std::min(threads, 3)) if (threads > 1)
    for (int v = 0; v < 3; v++) { ifft(vecs[v], c.dom); coset_fft(vecs[v], c.dom); }
    Fr zg = fr_pow_u64(c.dom.generator, m) - Fr::one();
    Fr zinv = zg.inv();
    for (size_t i = 0; i < m; i++) a[i] = (a[i] * b[i] - cc[i]) * zinv;
    coset_ifft(a.data(), c.dom);
    return a;
}

static void fr_vec_canonical(const Fr* v, size_t n, std::vector<uint64_t>& out) {
    out.resize(4 * n);
    for (size_t i = 0; i < n; i++) v[i].to_canonical(&out[4 * i]);
}

// create_proof (ark-groth16 0.3 prover.rs); z canonical (n x 4), r, s canonical
static void create_proof(const OracleCtx& c, const uint64_t* z_canon, const uint64_t* r, const uint64_t* s, uint8_t* out, int threads) {
    const size_t n = c.p + c.w;
    std::vector<Fr> z(n);
    for (size_t i = 0; i < n; i++) z[i] = Fr::from_canonical(z_canon + 4 * i);
    std::vector<Fr> h = witness_map(c, z.data(), threads);
    std::vector<uint64_t> h_canon;
    fr_vec_canonical(h.data(), h.size(), h_canon);
    const uint64_t* assignment = z_canon + 4;          // instance[1..] | witness
    const uint64_t* aux = z_canon + 4 * c.p;           // witness
    G1J h_acc, l_acc, a_acc, b1_acc;
    G2J b2_acc;
    bool r_zero = scalar_is_zero(r);
    // The five multi_scalar_mul calls of create_proof.  threads > 1 mirrors arkworks' `parallel` feature with ONE flat task
    // list over (MSM, window) - 5 x ~20 windows, the G2 windows (3x the cost) first - so that every host thread stays busy;
    // the window sums are combined exactly as the serial code does, so the result is the same group element.
    const size_t nh = std::min(c.h_query.size(), h.size()), nz = n - 1;
    const int ch = ark_window(nh), cl = ark_window(std::min<size_t>(c.l_query.size(), c.w)), cz = ark_window(nz);
    std::vector<G1J> wh(msm_num_windows(ch)), wl(msm_num_windows(cl)), wa(msm_num_windows(cz)), wb1(r_zero ? 0 : msm_num_windows(cz));
    std::vector<G2J> wb2(msm_num_windows(cz));
    struct Task { int msm, w; };
    std::vector<Task> tasks;
    for (size_t w = 0; w < wb2.size(); w++) tasks.push_back({4, (int)w});
    for (size_t w = 0; w < wh.size(); w++) tasks.push_back({0, (int)w});
    for (size_t w = 0; w < wl.size(); w++) tasks.push_back({1, (int)w});
    for (size_t w = 0; w < wa.size(); w++) tasks.push_back({2, (int)w});
    for (size_t w = 0; w < wb1.size(); w++) tasks.push_back({3, (int)w});
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) if (threads > 1)
    for (size_t t = 0; t < tasks.size(); t++) {
        const int w = tasks[t].w;
        switch (tasks[t].msm) {
            case 0: wh[w] = msm_window<Fq>(c.h_query.data(), h_canon.data(), nh, ch, w * ch); break;
            case 1: wl[w] = msm_window<Fq>(c.l_query.data(), aux, std::min<size_t>(c.l_query.size(), c.w), cl, w * cl); break;
            case 2: wa[w] = msm_window<Fq>(c.a_query.data() + 1, assignment, nz, cz, w * cz); break;
            case 3: wb1[w] = msm_window<Fq>(c.b_g1_query.data() + 1, assignment, nz, cz, w * cz); break;
            default: wb2[w] = msm_window<Fq2>(c.b_g2_query.data() + 1, assignment, nz, cz, w * cz); break;
        }
    }
    h_acc = msm_combine<Fq>(wh, ch);
    l_acc = msm_combine<Fq>(wl, cl);
    a_acc = msm_combine<Fq>(wa, cz);
    if (!r_zero) b1_acc = msm_combine<Fq>(wb1, cz);
    b2_acc = msm_combine<Fq2>(wb2, cz);
    // calculate_coeff(initial, query, vk_param, assignment) = initial + query[0] + acc + vk_param
    G1J delta1 = from_affine(c.delta_g1);
    G1J g_a = delta1.mul(r, 4);
    g_a.add_mixed(c.a_query[0]);
    g_a.add(a_acc);
    g_a.add_mixed(c.alpha_g1);
    G1J g1_b = G1J::identity();
    if (!r_zero) {
        g1_b = delta1.mul(s, 4);
        g1_b.add_mixed(c.b_g1_query[0]);
        g1_b.add(b1_acc);
        g1_b.add_mixed(c.beta_g1);
    }
    G2J g2_b = from_affine(c.delta_g2).mul(s, 4);
    g2_b.add_mixed(c.b_g2_query[0]);
    g2_b.add(b2_acc);
    g2_b.add_mixed(c.beta_g2);
    G1J g_c = g_a.mul(s, 4);
    g_c.add(g1_b.mul(r, 4));
    g_c.add(delta1.mul(r, 4).mul(s, 4).neg());
    g_c.add(l_acc);
    g_c.add(h_acc);
    write_g1_compressed(out, g_a.to_affine());
    write_g2_compressed(out + 48, g2_b.to_affine());
    write_g1_compressed(out + 144, g_c.to_affine());
}

static bool g_init = false;
static void ensure_init() {
    if (g_init) return;
    init_params<6>(g_fq, FQ_MOD, 381);
    init_params<4>(g_fr, FR_MOD, 255);
    g_init = true;
}

// ------------------------------------------------------------------------------------------------------------
// C interface (loaded with ctypes by the tests and the CPU-baseline leg of bench.py)
// ------------------------------------------------------------------------------------------------------------
extern "C" {

int oracle_max_threads() {
#ifdef _OPENMP
    return omp_get_num_procs();  // the host's cores (affinity-aware), NOT OMP_NUM_THREADS: torchrun exports OMP_NUM_THREADS=1
#else
    return 1;
#endif
}

// pk: `ProvingContext` bytes (groth16.rs:290-303).  CSR matrices with canonical coefficients.
void* oracle_ctx_create(const uint8_t* pk, size_t pk_len, uint64_t p, uint64_t w, uint64_t K, const uint64_t* const row_ptr[3],
                        const uint32_t* const col[3], const uint64_t* const coeff[3]) {
    ensure_init();
    std::unique_ptr<OracleCtx> c(new OracleCtx());
    size_t pos = 0;
    auto need = [&](size_t nb) { return nb <= pk_len - pos; };
    auto g1 = [&](G1A& o) { if (!need(96)) return false; o = read_g1_uncompressed(pk + pos); pos += 96; return true; };
    auto g2 = [&](G2A& o) { if (!need(192)) return false; o = read_g2_uncompressed(pk + pos); pos += 192; return true; };
    auto len = [&](uint64_t& o) { if (!need(8)) return false; memcpy(&o, pk + pos, 8); pos += 8; return true; };
    auto v1 = [&](std::vector<G1A>& v) {
        uint64_t cnt;
        if (!len(cnt) || cnt > (pk_len - pos) / 96) return false;
        v.resize(cnt);
        for (auto& e : v) g1(e);
        return true;
    };
    auto v2 = [&](std::vector<G2A>& v) {
        uint64_t cnt;
        if (!len(cnt) || cnt > (pk_len - pos) / 192) return false;
        v.resize(cnt);
        for (auto& e : v) g2(e);
        return true;
    };
    G2A gamma_g2;
    std::vector<G1A> gamma_abc;
    bool ok = g1(c->alpha_g1) && g2(c->beta_g2) && g2(gamma_g2) && g2(c->delta_g2) && v1(gamma_abc) && g1(c->beta_g1) &&
              g1(c->delta_g1) && v1(c->a_query) && v1(c->b_g1_query) && v2(c->b_g2_query) && v1(c->h_query) && v1(c->l_query);
    if (!ok || pos != pk_len) return nullptr;
    c->p = p; c->w = w; c->K = K;
    if (c->a_query.size() != p + w) return nullptr;
    for (int m = 0; m < 3; m++) {
        c->mat[m].row_ptr.assign(row_ptr[m], row_ptr[m] + K + 1);
        size_t nnz = row_ptr[m][K];
        c->mat[m].col.assign(col[m], col[m] + nnz);
        c->mat[m].coeff.resize(nnz);
        for (size_t e = 0; e < nnz; e++) c->mat[m].coeff[e] = Fr::from_canonical(coeff[m] + 4 * e);
    }
    c->dom = make_domain(K + p);
    return c.release();
}
void oracle_ctx_destroy(void* ctx) { delete (OracleCtx*)ctx; }
uint64_t oracle_ctx_domain_size(void* ctx) { return ((OracleCtx*)ctx)->dom.size; }

int oracle_prove(void* ctx, const uint64_t* z, const uint64_t* r, const uint64_t* s, uint8_t* out_proof, int threads) {
    if (!ctx) return 1;
    create_proof(*(OracleCtx*)ctx, z, r, s, out_proof, threads < 1 ? 1 : threads);
    return 0;
}

int oracle_witness_map(void* ctx, const uint64_t* z, uint64_t* out_h) {
    OracleCtx& c = *(OracleCtx*)ctx;
    size_t n = c.p + c.w;
    std::vector<Fr> zz(n);
    for (size_t i = 0; i < n; i++) zz[i] = Fr::from_canonical(z + 4 * i);
    std::vector<Fr> h = witness_map(c, zz.data(), 1);
    for (size_t i = 0; i < h.size(); i++) h[i].to_canonical(out_h + 4 * i);
    return 0;
}

int oracle_msm_g1(const uint8_t* bases, const uint64_t* scalars, size_t n, uint8_t* out, int threads) {
    ensure_init();
    std::vector<G1A> b(n);
    for (size_t i = 0; i < n; i++) b[i] = read_g1_uncompressed(bases + 96 * i);
    write_g1_uncompressed(out, msm<Fq>(b.data(), scalars, n, n, threads < 1 ? 1 : threads).to_affine());
    return 0;
}
int oracle_msm_g2(const uint8_t* bases, const uint64_t* scalars, size_t n, uint8_t* out, int threads) {
    ensure_init();
    std::vector<G2A> b(n);
    for (size_t i = 0; i < n; i++) b[i] = read_g2_uncompressed(bases + 192 * i);
    write_g2_uncompressed(out, msm<Fq2>(b.data(), scalars, n, n, threads < 1 ? 1 : threads).to_affine());
    return 0;
}
// out[i] = k_i * G for the standard generators (test helper)
int oracle_fixed_base(int group, const uint64_t* scalars, size_t n, uint8_t* out) {
    ensure_init();
    static const uint64_t G1X[6] = {0xfb3af00adb22c6bbULL, 0x6c55e83ff97a1aefULL, 0xa14e3a3f171bac58ULL, 0xc3688c4f9774b905ULL, 0x2695638c4fa9ac0fULL, 0x17f1d3a73197d794ULL};
    static const uint64_t G1Y[6] = {0x0caa232946c5e7e1ULL, 0xd03cc744a2888ae4ULL, 0x00db18cb2c04b3edULL, 0xfcf5e095d5d00af6ULL, 0xa09e30ed741d8ae4ULL, 0x08b3f481e3aaa0f1ULL};
    static const uint64_t G2X0[6] = {0xd48056c8c121bdb8ULL, 0x0bac0326a805bbefULL, 0xb4510b647ae3d177ULL, 0xc6e47ad4fa403b02ULL, 0x260805272dc51051ULL, 0x024aa2b2f08f0a91ULL};
    static const uint64_t G2X1[6] = {0xe5ac7d055d042b7eULL, 0x334cf11213945d57ULL, 0xb5da61bbdc7f5049ULL, 0x596bd0d09920b61aULL, 0x7dacd3a088274f65ULL, 0x13e02b6052719f60ULL};
    static const uint64_t G2Y0[6] = {0xe193548608b82801ULL, 0x923ac9cc3baca289ULL, 0x6d429a695160d12cULL, 0xadfd9baa8cbdd3a7ULL, 0x8cc9cdc6da2e351aULL, 0x0ce5d527727d6e11ULL};
    static const uint64_t G2Y1[6] = {0xaaa9075ff05f79beULL, 0x3f370d275cec1da1ULL, 0x267492ab572e99abULL, 0xcb3e287e85a763afULL, 0x32acd2b02bc28b99ULL, 0x0606c4a02ea734ccULL};
    if (group == 1) {
        G1J g = {Fq::from_canonical(G1X), Fq::from_canonical(G1Y), Fq::one()};
#pragma omp parallel for schedule(dynamic, 64)
        for (size_t i = 0; i < n; i++) write_g1_uncompressed(out + 96 * i, g.mul(scalars + 4 * i, 4).to_affine());
    } else {
        G2J g = {{Fq::from_canonical(G2X0), Fq::from_canonical(G2X1)}, {Fq::from_canonical(G2Y0), Fq::from_canonical(G2Y1)}, Fq2::one()};
#pragma omp parallel for schedule(dynamic, 64)
        for (size_t i = 0; i < n; i++) write_g2_uncompressed(out + 192 * i, g.mul(scalars + 4 * i, 4).to_affine());
    }
    return 0;
}
// in-place transform of canonical data: inverse / coset as in mp_ntt
int oracle_ntt(uint64_t* data, unsigned log_n, int inverse, int coset) {
    ensure_init();
    Domain d = make_domain((size_t)1 << log_n);
    std::vector<Fr> a(d.size);
    for (size_t i = 0; i < d.size; i++) a[i] = Fr::from_canonical(data + 4 * i);
    if (!inverse && !coset) fft(a.data(), d);
    else if (inverse && !coset) ifft(a.data(), d);
    else if (!inverse && coset) coset_fft(a.data(), d);
    else coset_ifft(a.data(), d);
    for (size_t i = 0; i < d.size; i++) a[i].to_canonical(data + 4 * i);
    return 0;
}
// element-wise field ops on canonical values: field 0 = Fq, 1 = Fr; op as mp_debug_field_op
int oracle_field_op(int field, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
    ensure_init();
    for (size_t i = 0; i < n; i++) {
        if (field == 0) {
            Fq x = Fq::from_canonical(a + 6 * i), y = b ? Fq::from_canonical(b + 6 * i) : Fq::zero(), r;
            switch (op) { case 0: r = x + y; break; case 1: r = x - y; break; case 2: r = x * y; break; case 3: r = x.sqr(); break; case 4: r = x.inv(); break; default: r = x.neg(); }
            r.to_canonical(out + 6 * i);
        } else {
            Fr x = Fr::from_canonical(a + 4 * i), y = b ? Fr::from_canonical(b + 4 * i) : Fr::zero(), r;
            switch (op) { case 0: r = x + y; break; case 1: r = x - y; break; case 2: r = x * y; break; case 3: r = x.sqr(); break; case 4: r = x.inv(); break; default: r = x.neg(); }
            r.to_canonical(out + 4 * i);
        }
    }
    return 0;
}

}  // extern "C"
