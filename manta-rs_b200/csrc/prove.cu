// Groth16 prover: device-resident proving context, batched proof pipeline and the C ABI around it.
//
// Replaces ark-groth16 0.3 `create_proof(circuit, pk, r, s)` as reached from
// manta-crypto/src/arkworks/groth16.rs:588-600 (`ProofSystem::prove`):
//
//   h      = witness_map(A, B, C, z)                                     (ntt.cu)
//   g_a    = r*delta_1 + a_query[0] + MSM(a_query[1..], z[1..]) + alpha_1
//   g1_b   = s*delta_1 + b_g1_query[0] + MSM(b_g1_query[1..], z[1..]) + beta_1
//   g2_b   = s*delta_2 + b_g2_query[0] + MSM(b_g2_query[1..], z[1..]) + beta_2
//   g_c    = s*g_a + r*g1_b - r*s*delta_1 + MSM(l_query, z[p..]) + MSM(h_query, h)
//   proof  = compress(g_a) | compress(g2_b) | compress(g_c)               (groth16.rs:184-195)
//
// Device formulation: the scalar vector is extended to z' = z | r | s | -rs | 1 and every query table gets four
// extra columns (delta / alpha / beta, or infinity where a term does not apply), so the fixed terms ride inside
// the MSMs (z_0 = 1 already selects query[0]).  A, B1, B2 and L then share ONE sorted digit list; H has its own.
// All tables hold the 16 window multiples 2^(16 t) P, so each MSM is a single 2^15-bucket problem with no Horner
// step.  Only s*g_a + r*g1_b remains as two 255-bit scalar multiplications in the finishing kernel.
#include <cstdlib>
#include <mutex>
#include <new>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "msm.cuh"
#include "ntt.cuh"

namespace mp {

constexpr int PROVE_C = 16;  // window bits of the circuit MSMs
constexpr int N_EXTRA = 4;   // r, s, -rs, 1

// Phases of a batch in stream order (CUDA-event intervals on the main stream).  The G2 MSM runs first because it needs
// only z; "msm_g2" = B-list sort + G2 bucket accumulation + the throughput part of its reduction.
enum Phase { PH_UPLOAD = 0, PH_PREP, PH_G2, PH_WITNESS_MAP, PH_SORT, PH_ACC_G1, PH_REDUCE, PH_FINISH, PH_COUNT };
static const char* const kPhaseNames[PH_COUNT] = {"upload",   "prep",              "msm_g2(sort+accumulate+reduce)", "witness_map(r1cs+ntt)",
                                                  "msm_sort", "msm_accumulate_g1", "msm_reduce_g1",                  "finish"};

// NVTX ranges around the enqueue of every phase, named like the timers of ark-groth16 0.3 `create_proof_with_reduction`
// ("Groth16::Prover" > "R1CS to QAP witness map", "Compute A", "Compute B in G1", "Compute B in G2", "Compute C", "Finish C"),
// so a timeline of this library lines up with the reference's `print-trace` output.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

}  // namespace mp

using namespace mp;

struct mp_ctx {
    int device = 0;
    uint64_t n = 0, p = 0, w = 0, K = 0, m = 0;
    unsigned log_m = 0;
    uint32_t zlen = 0;  // n + N_EXTRA
    R1csDev r1cs;
    NttDomain dom;
    MsmGeom gz{}, gh{};
    DevBuf tab_a, tab_b1, tab_l, tab_h, tab_b2;
    DevBuf valid_a, valid_b, valid_l, valid_h;  // per-table "base is not infinity" bitmaps
    size_t device_bytes = 0;
    bool glv = true;               // every A / B1 query point lies in the prime-order subgroup: the finishing kernel may use the GLV ladder
    std::mutex mu;
    std::mutex single_mu;          // serialises mp_prove callers on the cached one-proof batch
    mp_batch* single = nullptr;   // lazily created capacity-1 batch behind mp_prove
    // Batches of one context run their throughput kernels one after the other (co-running them on the same SMs costs
    // ~5 %); only the latency-bound tail of a batch (small weighted sums, finishing kernel) overlaps with the next one.
    cudaEvent_t last_heavy = nullptr;    // recorded by the batch enqueued last, after its last throughput kernel
    mp_batch* last_heavy_owner = nullptr;
};

struct mp_batch {
    mp_ctx* ctx = nullptr;
    size_t capacity = 0, count = 0;
    cudaStream_t st = nullptr, st2 = nullptr, st3 = nullptr;  // main stream; G2 path; A / L sorts of a small batch (beside the witness map)
    cudaEvent_t ev_sort_al = nullptr;
    cudaEvent_t ev_g2_heavy = nullptr, ev_g2 = nullptr, ev_heavy = nullptr, ev_tail_fork = nullptr, ev_sort_b = nullptr;
    cudaEvent_t ev_dom0 = nullptr, ev_dom1 = nullptr;  // around the dominant kernel (round-1 k_ba_bwd<Fq> of the G1 bucket trees)
    // slab pipeline of the tree levels (msm_impl.inc accumulate_ba): two highest-priority side streams for the inversion kernels
    cudaStream_t st_mid[2] = {nullptr, nullptr};
    cudaEvent_t ev_pipe[4] = {nullptr, nullptr, nullptr, nullptr};
    bool overlap = true;
    DevBuf z_canon, z_mont, rs, abc, s1, s2, h_canon;
    DevBuf sort_a_mem, sort_b_mem, sort_l_mem, sort_h_mem;
    MsmSortWs sort_a, sort_b, sort_l, sort_h;  // A | B1+B2 | L | H each get a list without their infinity bases
    DevBuf part_a, part_b1, part_l, part_h, part_b2;
    DevBuf pb_a, pb_b1, pb_l, pb_h, pb_b2, ba_mem_g1, ba_mem_g2;  // batched-affine point buffers and round scratch
    // Buffers whose lifetimes never overlap share memory in batches of more than 16 proofs, where the whole G2 MSM and the
    // witness map run on the main stream BEFORE the G1 MSMs: the G2 point buffer is dead once its row/column trees are done
    // and becomes the H point buffer; the G2 round scratch and the witness-map vectors live inside the G1 round scratch.
    // (Small batches run the G2 MSM on the second stream beside the G1 work and keep everything separate.)
    bool aliased = false;
    // Outer slabs of the MSM stage (MP_ACC_SLABS = 1, 2 or 4; 2 and 4 are also tried when the device runs out of memory): the
    // point buffers of the bucket trees - two thirds of a proof's device memory - are sized for capacity / acc_slabs proofs, and
    // bucket trees + row/column trees run slab after slab over them.  Everything per proof that outlives a slab (sorted lists,
    // row/column sums, results) keeps its full size; the kernels see a slab through offset pointers.
    size_t acc_slabs = 1, slab_cap = 0;
    void *p_pb_h = nullptr, *p_pb_b2 = nullptr, *p_abc = nullptr, *p_s1 = nullptr, *p_s2 = nullptr;
    MsmBaWs ba_g1, ba_g2;
    // one or two proofs: the bucket trees of A, B1, L start as soon as their lists are sorted (third stream, beside the witness
    // map), those of H follow the witness map on the main stream with a round scratch of their own
    DevBuf ba_mem_h, lad;
    MsmBaWs ba_h;
    // one or two proofs, key inside the prime-order subgroup: s * g_a and r * g1_b are computed as two more MSMs over the A and
    // B1 tables with the scalar vectors s z' and r z' (same launches as A, B1, L: no depth added) instead of two 128-step ladders
    // behind the A / B1 results.  zx: [3][count] rows of zlen scalars: r z' | z' | s z' (the A and B lists are sorted as 2 * count
    // vectors: one sort call each).
    DevBuf zx, red_sa, red_rb, ba_mem_abl;
    MsmBaWs ba_abl;
    cudaEvent_t ev_acc_abl = nullptr;
    MsmGeom gz_rc{}, gh_rc{};                                     // row/column stage of the bucket reduction
    DevBuf rc_a_mem, rc_b_mem, rc_b2_mem, rc_l_mem, rc_h_mem, pbrc_a, pbrc_b1, pbrc_l, pbrc_h, pbrc_b2, resrc_g1, resrc_g2;
    MsmSortWs rc_a, rc_b, rc_b2, rc_l, rc_h;  // B1 and B2 keep separate row/column lists: they may run on different streams
    bool use_ba = false;
    DevBuf res_g1, res_g2, red_a, red_b1, red_l, red_h, red_b2, proofs;
    DevBuf bad_dev;                // 1 + index of the first proof with a non-canonical scalar (k_prove_prep)
    uint32_t* bad_host = nullptr;  // pinned copy, read after the streams drain
    MsmGeom gz{}, gh{};
    cudaEvent_t ev[PH_COUNT + 1] = {};
    float phase_ms[PH_COUNT] = {};
    uint64_t launches = 0;
    bool ran = false, in_flight = false, upload_timed = false;
    bool abc_supplied = false;  // mp_prove_from_abc: the evaluation vectors are already in `abc`, skip the CSR products
    size_t device_bytes = 0;  // device memory behind this batch object (every per-proof buffer is allocated at creation)
};

namespace mp {

// ---------------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------------
// z' extras and the Montgomery copy of z used by the R1CS evaluation.  rs: [batch][2][8] canonical.
// Scalars cross the ABI canonical (< r): anything else would be recoded with a dropped carry and yield a wrong proof, so the
// first offending proof is reported through `bad` (1 + proof index; 0 = all fine) and the call fails with MP_ERR_INVALID_ARG.
MP_DEV bool fr_is_canonical(const Fr& x) {
    uint32_t t[Fr::N];
    return Fr::sub_raw(t, x.l, FrParams::mod()) != 0;  // borrow <=> x < r
}
__global__ void k_prove_prep(uint32_t* z_canon, uint32_t* z_mont, const uint32_t* __restrict__ rs, uint32_t n, uint32_t zlen, uint32_t* bad) {
    const uint32_t b = blockIdx.y;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t* zc = z_canon + (size_t)b * zlen * 8;
    uint32_t* zm = z_mont + (size_t)b * zlen * 8;
    if (i < n) {
        const Fr v = Fr::load(zc + (size_t)i * 8);
        if (!fr_is_canonical(v)) atomicMax(bad, b + 1);
        v.to_mont().store(zm + (size_t)i * 8);
    } else if (i == n) {
        Fr r = Fr::load(rs + (size_t)b * 16), s = Fr::load(rs + (size_t)b * 16 + 8);
        if (!fr_is_canonical(r) || !fr_is_canonical(s)) atomicMax(bad, b + 1);
        r.store(zc + (size_t)n * 8);
        s.store(zc + (size_t)(n + 1) * 8);
        // -(r s) canonical: mont(r) * s = r*s (canonical), then negate
        Fr rsv = r.to_mont() * s;
        rsv.neg().store(zc + (size_t)(n + 2) * 8);
        Fr one = Fr::zero();
        one.l[0] = 1;
        one.store(zc + (size_t)(n + 3) * 8);
    }
}

// zx rows: [r z' | z' | s z'], `cnt` rows each (z' = the extended scalar vector k_prove_prep completed): MSM_A(s z') = s g_a and
// MSM_B1(r z') = r g1_b when every table point has order r.  mont(r) * v is the canonical product for canonical v.
__global__ void k_prove_scale(const uint32_t* __restrict__ z_canon, const uint32_t* __restrict__ rs, uint32_t zlen, uint32_t cnt, uint32_t* zx) {
    const uint32_t b = blockIdx.y;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= zlen) return;
    const Fr v = Fr::load(z_canon + ((size_t)b * zlen + i) * 8);
    const Fr r = Fr::load(rs + (size_t)b * 16).to_mont(), s = Fr::load(rs + (size_t)b * 16 + 8).to_mont();
    (r * v).store(zx + ((size_t)b * zlen + i) * 8);
    v.store(zx + ((size_t)(cnt + b) * zlen + i) * 8);
    (s * v).store(zx + ((size_t)(2 * cnt + b) * zlen + i) * 8);
}

template <class F>
MP_COLD XYZZ<F> scalar_mul_affine(const Affine<F>& p, const uint32_t* k) {
    XYZZ<F> r = XYZZ<F>::inf();
    int bit = 254;
    while (bit >= 0 && !((k[bit >> 5] >> (bit & 31)) & 1)) bit--;
    for (; bit >= 0; bit--) {
        r = r.dbl();
        if ((k[bit >> 5] >> (bit & 31)) & 1) r = r.add_mixed_cold(p);
    }
    return r;
}

// ---------------------------------------------------------------------------------------------------------
// warp-cooperative scalar multiplication (the two 255-bit products s * g_a and r * g1_b of `create_proof`)
// ---------------------------------------------------------------------------------------------------------
// A warp instruction costs the multiplier pipe the same 4 cycles whether one lane or 32 are active, so a serial XYZZ
// ladder on one lane (the first cut: 2.8 ms per proof) wastes the pipe AND the latency.  Here the point lives in shared
// memory as 48-byte "slots" and every LEVEL of a formula - the field products that do not depend on each other - runs as
// ONE multiplication issued by the whole warp, each lane on its own pair of operands:
//   doubling (2008-s-1)  9 products -> 3 levels      addition (add-2008-s) 14 products -> 4 levels
// Level programs live in constant memory (per lane: two operand slots and a destination); the few sums and differences between
// levels are one-lane steps.  GLV (k = k1 + k2 lambda, 128-bit halves, table {P, phi P, P + phi P}) halves the doublings; the
// table stays in XYZZ because a full addition has the same depth as a mixed one once its products spread over the lanes.
// The formulas have no branches: an addition whose operands are equal (possible for chosen scalars) or whose accumulator
// is the point at infinity (only for points outside the prime-order subgroup) raises a flag, and the whole product is then
// redone by one lane with the complete XYZZ routines.
namespace coop {
constexpr int ACC = 0;      // slots 0..3: X, Y, ZZ, ZZZ of the accumulator
constexpr int TMP = 4;      // slots 4..19: temporaries
constexpr int TAB = 20;     // slots 20..31: table entries P, phi P, P + phi P (4 slots each)
constexpr int SLOTS = 32;
constexpr int REL = 64;     // operand indices >= REL address the selected table entry: slot = entry + (index - REL)
constexpr uint8_t NONE = 0xff;
constexpr int T(int k) { return TMP + k; }
// One level = up to 6 independent products; lane i multiplies slot a[i] by slot b[i] into slot d[i] (every lane runs the same
// code, operands are plain slots: the sums and differences a formula needs between levels are separate one-lane steps).
// post: slot subtracted from the product of a one-lane level (NONE = nothing).
struct Level { uint8_t n, post, a[6], b[6], d[6]; };
// doubling of the accumulator (dbl-2008-s-1, a = 0): U = 2Y, V = U^2, W = U V, S = X V, M = 3 X^2, X3 = M^2 - 2S,
// Y3 = M (S - X3) - W Y, ZZ3 = V ZZ, ZZZ3 = W ZZZ.   T7 = U, T8 = M, T9 = S - X3
__constant__ Level DBL1 = {4, NONE, {T(7), 0, T(7), T(7)}, {T(7), 0, 1, 3}, {T(0), T(1), T(2), T(3)}};                    // V, XX, U Y, U ZZZ
__constant__ Level DBL2 = {5, NONE, {0, T(8), T(0), T(0), T(0)}, {T(0), T(8), T(2), 2, T(3)}, {T(4), T(5), T(6), 2, 3}};  // S, M^2, W Y, ZZ3, ZZZ3
__constant__ Level DBL3 = {1, T(6), {T(8)}, {T(9)}, {1}};                                                                // Y3 = M (S - X3) - W Y
// accumulator + table entry E = (x2, y2, zz2, zzz2) (add-2008-s).   T13 = P = U2 - U1, T14 = R = S2 - S1, T15 = Q - X3
__constant__ Level ADD1 = {6, NONE, {0, REL + 0, 1, REL + 1, 2, 3}, {REL + 2, 2, REL + 3, 3, REL + 2, REL + 3},
                           {T(0), T(1), T(2), T(3), T(4), T(5)}};                                 // U1, U2, S1, S2, ZZ zz2, ZZZ zzz2
__constant__ Level ADD2 = {4, NONE, {T(13), T(14), T(13), T(13)}, {T(13), T(14), T(5), T(2)}, {T(6), T(7), T(8), T(9)}};   // PP, RR, P ZZZ zzz2, S1 P
__constant__ Level ADD3 = {5, NONE, {T(13), T(0), T(4), T(8), T(9)}, {T(6), T(6), T(6), T(6), T(6)}, {T(10), T(11), 2, 3, T(12)}};  // PPP, Q, ZZ3, ZZZ3, S1 PPP
__constant__ Level ADD4 = {1, T(12), {T(14)}, {T(15)}, {1}};                                                              // Y3 = R (Q - X3) - S1 PPP

MP_DEV Fq ld(const uint32_t* s, int i) { return Fq::load(s + 12 * i); }
MP_DEV void st(uint32_t* s, int i, const Fq& v) { v.store(s + 12 * i); }
MP_DEV int slot(uint8_t i, int entry) { return i >= REL ? entry + (i - REL) : i; }
// all loads of a level happen before any of its stores
MP_DEV void level(uint32_t* s, const Level& L, int entry, uint32_t lane) {
    Fq r;
    int dst = 0;
    if (lane < L.n) {
        r = ld(s, slot(L.a[lane], entry)) * ld(s, slot(L.b[lane], entry));
        if (L.post != NONE) r = r - ld(s, L.post);
        dst = L.d[lane];
    }
    __syncwarp();
    if (lane < L.n) st(s, dst, r);
    __syncwarp();
}
MP_DEV void dbl(uint32_t* s, uint32_t lane) {
    if (lane == 0) st(s, T(7), ld(s, 1).dbl());
    __syncwarp();
    level(s, DBL1, 0, lane);
    if (lane == 0) {
        const Fq xx = ld(s, T(1));
        st(s, T(8), xx.dbl() + xx);
    }
    __syncwarp();
    level(s, DBL2, 0, lane);
    if (lane == 0) {
        const Fq sv = ld(s, T(4)), x3 = ld(s, T(5)) - sv.dbl();
        st(s, 0, x3);
        st(s, T(9), sv - x3);
    }
    __syncwarp();
    level(s, DBL3, 0, lane);
}
// flag: raised when the formulas do not apply (accumulator at infinity, or equal x-coordinates: doubling / cancellation)
MP_DEV void add(uint32_t* s, int entry, uint32_t lane, uint32_t* flag) {
    level(s, ADD1, entry, lane);
    if (lane == 0) {
        const Fq p = ld(s, T(1)) - ld(s, T(0));
        if (p.is_zero() || ld(s, 2).is_zero()) *flag = 1;
        st(s, T(13), p);
    } else if (lane == 1) {
        st(s, T(14), ld(s, T(3)) - ld(s, T(2)));
    }
    __syncwarp();
    level(s, ADD2, entry, lane);
    level(s, ADD3, entry, lane);
    if (lane == 0) {
        const Fq q = ld(s, T(11)), x3 = ld(s, T(7)) - ld(s, T(10)) - q.dbl();
        st(s, 0, x3);
        st(s, T(15), q - x3);
    }
    __syncwarp();
    level(s, ADD4, entry, lane);
}
MP_DEV void copy_point(uint32_t* s, int dst, int src, uint32_t lane) {
    if (lane < 4) st(s, dst + lane, ld(s, src + lane));
    __syncwarp();
}
}  // namespace coop

// k = k2 lambda + k1 by long division (lambda = z^2 - 1, 128 bits; both halves < 2^128 because lambda^2 ~ r)
MP_COLD void glv_split(const uint32_t* k, uint32_t* k1, uint32_t* k2) {
    uint32_t rem[5] = {0, 0, 0, 0, 0}, q[8];
#pragma unroll
    for (int i = 0; i < 8; i++) q[i] = 0;
    for (int bit = 254; bit >= 0; bit--) {
#pragma unroll
        for (int i = 4; i > 0; i--) rem[i] = __funnelshift_l(rem[i - 1], rem[i], 1);
        rem[0] = (rem[0] << 1) | ((k[bit >> 5] >> (bit & 31)) & 1u);
        uint32_t t[5], borrow;
        sub_cc(t[0], rem[0], FR_GLV_LAMBDA[0]);
        subc_cc(t[1], rem[1], FR_GLV_LAMBDA[1]);
        subc_cc(t[2], rem[2], FR_GLV_LAMBDA[2]);
        subc_cc(t[3], rem[3], FR_GLV_LAMBDA[3]);
        subc_cc(t[4], rem[4], 0);
        subc(borrow, 0, 0);
        if (!borrow) {
#pragma unroll
            for (int i = 0; i < 5; i++) rem[i] = t[i];
            q[bit >> 5] |= 1u << (bit & 31);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) { k1[i] = rem[i]; k2[i] = q[i]; }
}

// k * P by ONE warp; P given in XYZZ; the result lands in out (XYZZ, 48 words).  s: the warp's coop::SLOTS slots; sc: 16 words
// of scratch (k1 | k2 | flag).  glv = 0 runs the plain 255-bit ladder (keys with points outside the prime-order subgroup:
// ark's double-and-add is defined there, phi(P) = lambda P is not).
MP_DEV void warp_scalar_mul(uint32_t* s, uint32_t* sc, const XYZZ<Fq>& p, const uint32_t* k, int glv, uint32_t lane, uint32_t* out) {
    using namespace coop;
    uint32_t* flag = sc + 8;
    bool kz = true;
#pragma unroll
    for (int i = 0; i < 8; i++) kz = kz && k[i] == 0;
    if (p.is_inf() || kz) {
        if (lane == 0) XYZZ<Fq>::inf().store(out);
        __syncwarp();
        return;
    }
    if (lane == 0) {
        *flag = 0;
        if (glv) glv_split(k, sc, sc + 4);
        else
            for (int i = 0; i < 8; i++) sc[i] = k[i];
        st(s, TAB + 0, p.X); st(s, TAB + 1, p.Y); st(s, TAB + 2, p.ZZ); st(s, TAB + 3, p.ZZZ);
        if (glv) {  // phi(P) = (beta x, y): scale X
            st(s, TAB + 4, p.X.mul_cold(Fq::from_const(FQ_GLV_BETA))); st(s, TAB + 5, p.Y); st(s, TAB + 6, p.ZZ); st(s, TAB + 7, p.ZZZ);
        }
    }
    __syncwarp();
    int bit;
    if (glv) {
        copy_point(s, ACC, TAB, lane);
        add(s, TAB + 4, lane, flag);           // P + phi P (never degenerate: phi P = lambda P, lambda != +-1)
        copy_point(s, TAB + 8, ACC, lane);
        const uint32_t* k1 = sc;
        const uint32_t* k2 = sc + 4;
        bit = 127;
        while (bit >= 0 && !(((k1[bit >> 5] | k2[bit >> 5]) >> (bit & 31)) & 1)) bit--;
        bool first = true;
        for (; bit >= 0; bit--) {
            const uint32_t sel = ((k1[bit >> 5] >> (bit & 31)) & 1) | (((k2[bit >> 5] >> (bit & 31)) & 1) << 1);
            if (!first) dbl(s, lane);
            if (sel) {
                const int entry = TAB + 4 * (int)(sel - 1);   // 1: P, 2: phi P, 3: P + phi P
                if (first) copy_point(s, ACC, entry, lane);
                else add(s, entry, lane, flag);
                first = false;
            }
        }
    } else {
        bit = 254;
        while (bit >= 0 && !((sc[bit >> 5] >> (bit & 31)) & 1)) bit--;
        copy_point(s, ACC, TAB, lane);
        for (bit--; bit >= 0; bit--) {
            dbl(s, lane);
            if ((sc[bit >> 5] >> (bit & 31)) & 1) add(s, TAB, lane, flag);
        }
    }
    __syncwarp();
    if (lane == 0) {
        XYZZ<Fq> r;
        if (*flag) {  // a degenerate addition was met: complete formulas, one lane
            r = XYZZ<Fq>::inf();
            int b = 254;
            while (b >= 0 && !((k[b >> 5] >> (b & 31)) & 1)) b--;
            for (; b >= 0; b--) {
                r = r.dbl();
                if ((k[b >> 5] >> (b & 31)) & 1) r = r.add(p);
            }
        } else {
            r = {ld(s, 0), ld(s, 1), ld(s, 2), ld(s, 3)};
        }
        r.store(out);
    }
    __syncwarp();
}

MP_DEV void write_fq_le(uint8_t* out, const Fq& canon) {
#pragma unroll
    for (int i = 0; i < 12; i++) {
        uint32_t v = canon.l[i];
        out[4 * i] = (uint8_t)v;
        out[4 * i + 1] = (uint8_t)(v >> 8);
        out[4 * i + 2] = (uint8_t)(v >> 16);
        out[4 * i + 3] = (uint8_t)(v >> 24);
    }
}

// ark-serialize compressed forms
MP_COLD void compress_g1(uint8_t* out, const Affine<Fq>& p) {
    if (p.is_inf()) {
        for (int i = 0; i < 48; i++) out[i] = 0;
        out[47] = 0x40;
        return;
    }
    write_fq_le(out, p.x.from_mont());
    if (Fq::canonical_gt_half(p.y.from_mont())) out[47] |= 0x80;
}
MP_COLD void compress_g2(uint8_t* out, const Affine<Fq2>& p) {
    if (p.is_inf()) {
        for (int i = 0; i < 96; i++) out[i] = 0;
        out[95] = 0x40;
        return;
    }
    write_fq_le(out, p.x.c0.from_mont());
    write_fq_le(out + 48, p.x.c1.from_mont());
    Fq y1 = p.y.c1.from_mont();
    bool larger = y1.is_zero() ? Fq::canonical_gt_half(p.y.c0.from_mont()) : Fq::canonical_gt_half(y1);
    if (larger) out[95] |= 0x80;
}

// G1 half of the proof, one block per proof, four warps:
//   warp 0: s * g_a          warp 1: r * g1_b          (warp-cooperative ladders)
//   warp 2: L + H            warp 3: g_a -> affine -> bytes       (one lane each)
// then thread 0 assembles g_c = s g_a + r g1_b + L + H (the -rs delta_1 term rides inside the L MSM) and writes its bytes.
// The G2 element is finished by k_prove_finish_g2 on the stream of the G2 MSM, so that for a single proof the G2 pipeline
// (three Fq products deep per multiplication) is no longer in front of the ladders on the critical path.
// res_g1: [4][batch] XYZZ (A, B1, L, H); res_g2: [batch] XYZZ.
constexpr int FINISH_THREADS = 128;
__global__ void __launch_bounds__(FINISH_THREADS) k_prove_finish(const XYZZ<Fq>* __restrict__ res_g1, const uint32_t* __restrict__ rs,
                                                                uint32_t batch, int glv, uint8_t* proofs) {
    __shared__ __align__(16) uint32_t slots[2][coop::SLOTS * 12];
    __shared__ __align__(16) uint32_t scratch[2][16];
    __shared__ __align__(16) uint32_t sh_sa[48], sh_rb[48], sh_lh[48];
    const uint32_t b = blockIdx.x;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* out = proofs + (size_t)b * MP_PROOF_BYTES;
    const uint32_t* r = rs + (size_t)b * 16;
    const uint32_t* s = r + 8;
    if (warp == 0) {
        warp_scalar_mul(slots[0], scratch[0], XYZZ<Fq>::load(res_g1 + b), s, glv, lane, sh_sa);
    } else if (warp == 1) {
        warp_scalar_mul(slots[1], scratch[1], XYZZ<Fq>::load(res_g1 + (size_t)batch + b), r, glv, lane, sh_rb);
    } else if (lane == 0) {
        if (warp == 2) {
            XYZZ<Fq> lh = XYZZ<Fq>::load(res_g1 + (size_t)2 * batch + b).add(XYZZ<Fq>::load(res_g1 + (size_t)3 * batch + b));
            lh.store(sh_lh);
        } else {
            compress_g1(out, XYZZ<Fq>::load(res_g1 + b).to_affine());
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        XYZZ<Fq> c = XYZZ<Fq>::load(sh_sa).add(XYZZ<Fq>::load(sh_rb)).add(XYZZ<Fq>::load(sh_lh));
        compress_g1(out + 144, c.to_affine());
    }
}
// The same in two kernels, for one or two proofs whose A / B1 / L pipeline runs on its own stream ahead of the H pipeline:
// the ladders (and the bytes of g_a) as soon as the A and B1 results exist, the assembly of g_c once L and H are there.
// lad: [batch][2] XYZZ (s g_a, r g1_b).
__global__ void __launch_bounds__(96) k_prove_ladders(const XYZZ<Fq>* __restrict__ res_g1, const uint32_t* __restrict__ rs, uint32_t batch, int glv,
                                                     uint32_t* lad, uint8_t* proofs) {
    __shared__ __align__(16) uint32_t slots[2][coop::SLOTS * 12];
    __shared__ __align__(16) uint32_t scratch[2][16];
    const uint32_t b = blockIdx.x;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t* r = rs + (size_t)b * 16;
    const uint32_t* s = r + 8;
    if (warp == 0) warp_scalar_mul(slots[0], scratch[0], XYZZ<Fq>::load(res_g1 + b), s, glv, lane, lad + (size_t)b * 96);
    else if (warp == 1) warp_scalar_mul(slots[1], scratch[1], XYZZ<Fq>::load(res_g1 + (size_t)batch + b), r, glv, lane, lad + (size_t)b * 96 + 48);
    else if (lane == 0) compress_g1(proofs + (size_t)b * MP_PROOF_BYTES, XYZZ<Fq>::load(res_g1 + b).to_affine());
}
__global__ void __launch_bounds__(32) k_prove_assemble(const XYZZ<Fq>* __restrict__ res_g1, const uint32_t* __restrict__ lad, uint32_t batch,
                                                      uint8_t* proofs) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    XYZZ<Fq> c = XYZZ<Fq>::load(lad + (size_t)b * 96).add(XYZZ<Fq>::load(lad + (size_t)b * 96 + 48));
    c = c.add(XYZZ<Fq>::load(res_g1 + (size_t)2 * batch + b)).add(XYZZ<Fq>::load(res_g1 + (size_t)3 * batch + b));
    compress_g1(proofs + (size_t)b * MP_PROOF_BYTES + 144, c.to_affine());
}
// The finishing step when s g_a and r g1_b came out of MSMs (res_g1: [6][batch] XYZZ: A, B1, L, H, s g_a, r g1_b): one block
// per proof; warp 0 writes the bytes of g_a, warp 1 those of g_c = s g_a + r g1_b + L + H.
__global__ void __launch_bounds__(64) k_prove_assemble_msm(const XYZZ<Fq>* __restrict__ res_g1, uint32_t batch, uint8_t* proofs) {
    const uint32_t b = blockIdx.x;
    if (threadIdx.x & 31) return;
    uint8_t* out = proofs + (size_t)b * MP_PROOF_BYTES;
    if (threadIdx.x == 0) {
        compress_g1(out, XYZZ<Fq>::load(res_g1 + b).to_affine());
    } else {
        XYZZ<Fq> c = XYZZ<Fq>::load(res_g1 + (size_t)4 * batch + b).add(XYZZ<Fq>::load(res_g1 + (size_t)5 * batch + b));
        c = c.add(XYZZ<Fq>::load(res_g1 + (size_t)2 * batch + b)).add(XYZZ<Fq>::load(res_g1 + (size_t)3 * batch + b));
        compress_g1(out + 144, c.to_affine());
    }
}
// g2_b -> affine -> the middle 96 bytes of the proof; one thread per proof
__global__ void __launch_bounds__(32) k_prove_finish_g2(const XYZZ<Fq2>* __restrict__ res_g2, uint32_t batch, uint8_t* proofs) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < batch) compress_g2(proofs + (size_t)b * MP_PROOF_BYTES + 48, XYZZ<Fq2>::load(res_g2 + b).to_affine());
}

// [r] P == infinity for every point of a query (prime-order subgroup membership; enables the GLV ladder of the finishing kernel)
template <class F>
__global__ void __launch_bounds__(64) k_subgroup_check(const uint32_t* __restrict__ pts, size_t n, uint32_t* outside) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    Affine<F> p = Affine<F>::load(pts + i * Affine<F>::WORDS);
    if (p.is_inf()) return;
    XYZZ<F> acc = XYZZ<F>::inf();
    for (int bit = 254; bit >= 0; bit--) {
        acc = acc.dbl();
        if ((FR_MOD[bit >> 5] >> (bit & 31)) & 1) acc = acc.add_mixed_cold(p);
    }
    if (!acc.is_inf()) atomicAdd(outside, 1u);
}

// ---------------------------------------------------------------------------------------------------------
// host helpers
// ---------------------------------------------------------------------------------------------------------
// Assemble `stride` affine points on the device: [front_pad x inf | query (len) | inf ... | extras at n..n+3]
template <bool G2>
static int assemble_bases(DevBuf& out, size_t stride, size_t front_pad, const uint8_t* query, size_t len, size_t n,
                          const uint8_t* const extras[N_EXTRA], cudaStream_t st) {
    const size_t pb = G2 ? MP_G2_BYTES : MP_G1_BYTES;
    MP_TRY(out.alloc(stride * pb));
    // ark bytes for infinity: all zero with 0x40 in the last byte -> fill via host staging
    std::vector<uint8_t> host(stride * pb, 0);
    for (size_t i = 0; i < stride; i++) host[i * pb + pb - 1] = 0x40;
    if (len) memcpy(host.data() + front_pad * pb, query, len * pb);
    for (int e = 0; e < N_EXTRA; e++)
        if (extras && extras[e]) memcpy(host.data() + (n + e) * pb, extras[e], pb);
    MP_CUDA_TRY(cudaMemcpyAsync(out.p, host.data(), stride * pb, cudaMemcpyHostToDevice, st));
    MP_CUDA_TRY(cudaStreamSynchronize(st));
    if (G2) MP_TRY(points_from_ark_g2(out.p, out.p, stride, st));
    else MP_TRY(points_from_ark_g1(out.p, out.p, stride, st));
    return MP_OK;
}

// The GLV ladder of the finishing kernel is only valid on the prime-order subgroup; the key is loaded like the reference's
// `deserialize_unchecked`, so membership of the points behind g_a and g1_b is established here, once per context.
static int check_subgroup_g1(mp_ctx* c, const DevBuf& bases, size_t n, cudaStream_t st) {
    DevBuf cnt;
    MP_TRY(cnt.alloc(4));
    MP_CUDA_TRY(cudaMemsetAsync(cnt.p, 0, 4, st));
    k_subgroup_check<Fq><<<div_up(n, 64), 64, 0, st>>>(bases.as<uint32_t>(), n, cnt.as<uint32_t>());
    MP_KERNEL_CHECK();
    uint32_t outside = 0;
    MP_CUDA_TRY(cudaMemcpyAsync(&outside, cnt.p, 4, cudaMemcpyDeviceToHost, st));
    MP_CUDA_TRY(cudaStreamSynchronize(st));
    if (outside) c->glv = false;
    return MP_OK;
}

template <bool G2>
static int build_table(mp_ctx* c, const MsmGeom& g, DevBuf& table, DevBuf& bases, DevBuf& valid, bool valid_accumulate, cudaStream_t st) {
    const size_t pb = G2 ? MP_G2_BYTES : MP_G1_BYTES;
    MP_TRY(table.alloc((size_t)g.rows * g.table_stride * pb));
    c->device_bytes += table.bytes;
    if (!valid_accumulate) MP_TRY(valid.alloc(((size_t)g.table_stride + 31) / 32 * 4));
    if (G2) MP_TRY(msm_validity_g2(bases.p, g.table_stride, valid.as<uint32_t>(), valid_accumulate, st));
    else MP_TRY(msm_validity_g1(bases.p, g.table_stride, valid.as<uint32_t>(), valid_accumulate, st));
    if (G2) MP_TRY(msm_build_table_g2(g, bases.p, g.table_stride, table.p, st));
    else MP_TRY(msm_build_table_g1(g, bases.p, g.table_stride, table.p, st));
    MP_CUDA_TRY(cudaStreamSynchronize(st));
    bases.release();
    return MP_OK;
}

static int ctx_create_impl(const mp_pk_view* pk, const mp_r1cs_view* r1cs, int device, mp_ctx* c) {
    MP_TRY(use_device(device));
    c->device = device;
    c->p = r1cs->num_instance;
    c->w = r1cs->num_witness;
    c->K = r1cs->num_constraints;
    c->n = c->p + c->w;
    if (c->p == 0 || c->n >= (1u << 26)) { set_error_detail("bad R1CS shape"); return MP_ERR_INVALID_ARG; }
    if (pk->a_len != c->n || pk->b_g1_len != c->n || pk->b_g2_len != c->n || pk->l_len != c->w) {
        set_error_detail("proving key does not match the R1CS shape (n=%llu, w=%llu; a=%llu b1=%llu b2=%llu l=%llu)",
                         (unsigned long long)c->n, (unsigned long long)c->w, (unsigned long long)pk->a_len,
                         (unsigned long long)pk->b_g1_len, (unsigned long long)pk->b_g2_len, (unsigned long long)pk->l_len);
        return MP_ERR_INVALID_ARG;
    }
    c->log_m = 0;
    while (((uint64_t)1 << c->log_m) < c->K + c->p) c->log_m++;
    c->m = (uint64_t)1 << c->log_m;
    if (c->log_m > 20) { set_error_detail("prover: domain 2^%u exceeds 2^20 (the stand-alone transform mp_ntt goes to 2^26)", c->log_m); return MP_ERR_UNSUPPORTED; }
    if (pk->h_len + 1 < c->m) { set_error_detail("h_query shorter than m - 1"); return MP_ERR_INVALID_ARG; }
    c->zlen = (uint32_t)c->n + N_EXTRA;
    cudaStream_t st = 0;
    MP_TRY(r1cs_upload(c->r1cs, r1cs, st));
    MP_TRY(ntt_domain_create(c->dom, c->log_m, st));
    c->gz = msm_geom(PROVE_C, 1, c->zlen, c->zlen, 1);
    c->gh = msm_geom(PROVE_C, 1, (uint32_t)c->m, (uint32_t)c->m, 1);
    {
        DevBuf bases;
        const uint8_t* ex_a[N_EXTRA] = {pk->delta_g1, nullptr, nullptr, pk->alpha_g1};
        MP_TRY(assemble_bases<false>(bases, c->zlen, 0, pk->a_query, c->n, c->n, ex_a, st));
        MP_TRY(check_subgroup_g1(c, bases, c->zlen, st));
        MP_TRY(build_table<false>(c, c->gz, c->tab_a, bases, c->valid_a, false, st));
        const uint8_t* ex_b1[N_EXTRA] = {nullptr, pk->delta_g1, nullptr, pk->beta_g1};
        MP_TRY(assemble_bases<false>(bases, c->zlen, 0, pk->b_g1_query, c->n, c->n, ex_b1, st));
        MP_TRY(check_subgroup_g1(c, bases, c->zlen, st));
        MP_TRY(build_table<false>(c, c->gz, c->tab_b1, bases, c->valid_b, false, st));
        const uint8_t* ex_l[N_EXTRA] = {nullptr, nullptr, pk->delta_g1, nullptr};
        MP_TRY(assemble_bases<false>(bases, c->zlen, c->p, pk->l_query, c->w, c->n, ex_l, st));
        MP_TRY(build_table<false>(c, c->gz, c->tab_l, bases, c->valid_l, false, st));
        size_t hl = std::min<uint64_t>(pk->h_len, c->m);
        MP_TRY(assemble_bases<false>(bases, c->m, 0, pk->h_query, hl, c->m, nullptr, st));
        MP_TRY(build_table<false>(c, c->gh, c->tab_h, bases, c->valid_h, false, st));
        const uint8_t* ex_b2[N_EXTRA] = {nullptr, pk->delta_g2, nullptr, pk->beta_g2};
        MP_TRY(assemble_bases<true>(bases, c->zlen, 0, pk->b_g2_query, c->n, c->n, ex_b2, st));
        MP_TRY(build_table<true>(c, c->gz, c->tab_b2, bases, c->valid_b, true, st));  // B list keeps a scalar if either B1 or B2 base is finite
    }
    MP_CUDA_TRY(cudaDeviceSynchronize());
    if (const char* e = getenv("MP_NO_GLV"))
        if (e[0] && e[0] != '0') c->glv = false;  // test hook: plain 255-bit ladder
    return MP_OK;
}

static int batch_create_impl(mp_ctx* c, size_t cap, int high_priority, size_t acc_slabs, mp_batch* b) {
    MP_TRY(use_device(c->device));
    b->ctx = c;
    b->capacity = cap;
    b->acc_slabs = (cap > 16 && msm_use_batched_affine()) ? acc_slabs : 1;
    b->slab_cap = (cap + b->acc_slabs - 1) / b->acc_slabs;
    const uint64_t alloc0 = dev_alloc_counter();
    int prio_least = 0, prio_greatest = 0;
    MP_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    const int prio = high_priority ? prio_greatest : prio_least;
    MP_CUDA_TRY(cudaStreamCreateWithPriority(&b->st, cudaStreamNonBlocking, prio));
    MP_CUDA_TRY(cudaStreamCreateWithPriority(&b->st2, cudaStreamNonBlocking, prio));
    MP_CUDA_TRY(cudaStreamCreateWithPriority(&b->st3, cudaStreamNonBlocking, prio));
    MP_CUDA_TRY(cudaEventCreateWithFlags(&b->ev_sort_al, cudaEventDisableTiming));
    MP_CUDA_TRY(cudaEventCreateWithFlags(&b->ev_acc_abl, cudaEventDisableTiming));
    for (auto& e : b->ev) MP_CUDA_TRY(cudaEventCreate(&e));
    for (cudaEvent_t* e : {&b->ev_g2_heavy, &b->ev_g2, &b->ev_tail_fork, &b->ev_sort_b, &b->ev_dom0, &b->ev_dom1}) MP_CUDA_TRY(cudaEventCreate(e));
    MP_CUDA_TRY(cudaEventCreateWithFlags(&b->ev_heavy, cudaEventDisableTiming));
    const size_t m = c->m;
    MP_TRY(b->z_canon.alloc(cap * c->zlen * 32));
    MP_TRY(b->z_mont.alloc(cap * c->zlen * 32));
    MP_TRY(b->rs.alloc(cap * 64));
    b->use_ba = msm_use_batched_affine();
    b->aliased = b->use_ba && cap > 16;
    if (!b->aliased) {
        MP_TRY(b->abc.alloc(cap * 3 * m * 32));
        MP_TRY(b->s1.alloc(cap * 3 * m * 32));
        MP_TRY(b->s2.alloc(cap * 3 * m * 32));
        b->p_abc = b->abc.p; b->p_s1 = b->s1.p; b->p_s2 = b->s2.p;
    }
    MP_TRY(b->h_canon.alloc(cap * m * 32));
    b->gz = msm_geom(PROVE_C, 1, c->zlen, c->zlen, cap * 2);  // G1 launches carry 4 jobs, the G2 launch one
    b->gh = msm_geom(PROVE_C, 1, (uint32_t)c->m, (uint32_t)c->m, cap * 2);
    // MP_LADDERS_AS_MSM=1 (measured alternative, off by default): the A and B lists also carry the scaled vectors (see `zx`)
    bool ext = false;
    if (const char* e = getenv("MP_LADDERS_AS_MSM")) ext = e[0] == '1' && msm_use_batched_affine() && cap <= 2;
    MP_TRY(msm_sort_ws_alloc(b->sort_a, b->gz, ext ? 2 * cap : cap, b->sort_a_mem));
    MP_TRY(msm_sort_ws_alloc(b->sort_b, b->gz, ext ? 2 * cap : cap, b->sort_b_mem));
    MP_TRY(msm_sort_ws_alloc(b->sort_l, b->gz, cap, b->sort_l_mem));
    MP_TRY(msm_sort_ws_alloc(b->sort_h, b->gh, cap, b->sort_h_mem));
    const size_t g1w = XYZZ<Fq>::WORDS * 4, g2w = XYZZ<Fq2>::WORDS * 4;
    if (b->use_ba) {
        const size_t scap = b->slab_cap;
        MP_TRY(b->pb_a.alloc((ext ? 2 : 1) * scap * b->gz.p_cap * MP_G1_BYTES));
        MP_TRY(b->pb_b1.alloc((ext ? 2 : 1) * scap * b->gz.p_cap * MP_G1_BYTES));
        MP_TRY(b->pb_l.alloc(scap * b->gz.p_cap * MP_G1_BYTES));
        const MsmGeom geoms_g1[4] = {b->gz, b->gz, b->gz, b->gh};
        const size_t pbh = scap * b->gh.p_cap * MP_G1_BYTES, pbb2 = scap * b->gz.p_cap * MP_G2_BYTES;
        const size_t ba1 = msm_ba_ws_bytes(geoms_g1, 4, cap, false), ba2 = msm_ba_ws_bytes(&b->gz, 1, cap, true);
        const size_t wm = (cap * 3 * m * 32 + 255) & ~size_t(255);
        if (b->aliased) {
            MP_TRY(b->pb_h.alloc(std::max(pbh, pbb2)));
            b->p_pb_h = b->p_pb_b2 = b->pb_h.p;
            MP_TRY(b->ba_mem_g1.alloc(std::max(std::max(ba1, ba2), 3 * wm)));
            // the G2 rounds borrow the (larger) G1 round scratch: give them the largest slab of proofs that fits - fewer, larger
            // launches per tree level
            size_t g2_slab = cap;
            while (g2_slab > msm_ba_slab(cap, true) && msm_ba_ws_bytes(&b->gz, 1, cap, true, g2_slab) > b->ba_mem_g1.bytes) g2_slab = (g2_slab + 1) / 2;
            if (msm_ba_ws_bytes(&b->gz, 1, cap, true, g2_slab) > b->ba_mem_g1.bytes) g2_slab = 0;
            if (const char* e = getenv("MP_G2_SLAB_DEFAULT"))
                if (e[0] == '1') g2_slab = 0;   // A/B hook
            msm_ba_ws_bind(b->ba_g2, &b->gz, 1, cap, true, b->ba_mem_g1.p, g2_slab);
            b->p_abc = b->ba_mem_g1.p;
            b->p_s1 = b->ba_mem_g1.as<char>() + wm;
            b->p_s2 = b->ba_mem_g1.as<char>() + 2 * wm;
        } else {
            MP_TRY(b->pb_h.alloc(pbh));
            MP_TRY(b->pb_b2.alloc(pbb2));
            b->p_pb_h = b->pb_h.p;
            b->p_pb_b2 = b->pb_b2.p;
            MP_TRY(b->ba_mem_g1.alloc(ba1));
            MP_TRY(b->ba_mem_g2.alloc(ba2));
            msm_ba_ws_bind(b->ba_g2, &b->gz, 1, cap, true, b->ba_mem_g2.p);
        }
        msm_ba_ws_bind(b->ba_g1, geoms_g1, 4, cap, false, b->ba_mem_g1.p);
        if (cap <= 2) {
            MP_TRY(b->ba_mem_h.alloc(msm_ba_ws_bytes(&b->gh, 1, cap, false)));
            msm_ba_ws_bind(b->ba_h, &b->gh, 1, cap, false, b->ba_mem_h.p);
            MP_TRY(b->lad.alloc(cap * 2 * XYZZ<Fq>::WORDS * 4));
            if (ext) {
                const MsmGeom geoms_abl[5] = {b->gz, b->gz, b->gz, b->gz, b->gz};
                MP_TRY(b->ba_mem_abl.alloc(msm_ba_ws_bytes(geoms_abl, 5, cap, false)));
                msm_ba_ws_bind(b->ba_abl, geoms_abl, 5, cap, false, b->ba_mem_abl.p);
                MP_TRY(b->zx.alloc(3 * cap * c->zlen * 32));
                MP_TRY(b->red_sa.alloc(msm_reduce_scratch_bytes(b->gz, cap, false)));
                MP_TRY(b->red_rb.alloc(msm_reduce_scratch_bytes(b->gz, cap, false)));
            }
        }
        b->ba_g1.ev_bwd0 = b->ev_dom0;
        b->ba_g1.ev_bwd1 = b->ev_dom1;
        if (b->ba_g1.n_sets == 2 || b->ba_g2.n_sets == 2) {
            for (auto& s : b->st_mid) MP_CUDA_TRY(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, prio_greatest));
            for (auto& e : b->ev_pipe) MP_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            for (MsmBaWs* w : {&b->ba_g1, &b->ba_g2}) {   // both run on the main stream of a large batch, one after the other
                w->mid_st[0] = b->st_mid[0]; w->mid_st[1] = b->st_mid[1];
                w->ev_fwd[0] = b->ev_pipe[0]; w->ev_fwd[1] = b->ev_pipe[1];
                w->ev_mid[0] = b->ev_pipe[2]; w->ev_mid[1] = b->ev_pipe[3];
            }
        }
        b->gz_rc = msm_geom_rc(b->gz);
        b->gh_rc = msm_geom_rc(b->gh);
        MP_TRY(msm_sort_ws_alloc(b->rc_a, b->gz_rc, cap, b->rc_a_mem));
        MP_TRY(msm_sort_ws_alloc(b->rc_b, b->gz_rc, cap, b->rc_b_mem));
        MP_TRY(msm_sort_ws_alloc(b->rc_b2, b->gz_rc, cap, b->rc_b2_mem));
        MP_TRY(msm_sort_ws_alloc(b->rc_l, b->gz_rc, cap, b->rc_l_mem));
        MP_TRY(msm_sort_ws_alloc(b->rc_h, b->gh_rc, cap, b->rc_h_mem));
        MP_TRY(b->pbrc_a.alloc(cap * b->gz_rc.p_cap * MP_G1_BYTES));
        MP_TRY(b->pbrc_b1.alloc(cap * b->gz_rc.p_cap * MP_G1_BYTES));
        MP_TRY(b->pbrc_l.alloc(cap * b->gz_rc.p_cap * MP_G1_BYTES));
        MP_TRY(b->pbrc_h.alloc(cap * b->gh_rc.p_cap * MP_G1_BYTES));
        MP_TRY(b->pbrc_b2.alloc(cap * b->gz_rc.p_cap * MP_G2_BYTES));
        MP_TRY(b->resrc_g1.alloc(4 * cap * 2 * g1w));
        MP_TRY(b->resrc_g2.alloc(cap * 2 * g2w));
    } else {
        MP_TRY(b->part_a.alloc(cap * b->gz.max_items * g1w));
        MP_TRY(b->part_b1.alloc(cap * b->gz.max_items * g1w));
        MP_TRY(b->part_l.alloc(cap * b->gz.max_items * g1w));
        MP_TRY(b->part_h.alloc(cap * b->gh.max_items * g1w));
        MP_TRY(b->part_b2.alloc(cap * b->gz.max_items * g2w));
    }
    MP_TRY(b->res_g1.alloc(6 * cap * g1w));   // A, B1, L, H (+ s g_a, r g1_b as MSMs for one or two proofs)
    MP_TRY(b->res_g2.alloc(cap * g2w));
    MP_TRY(b->red_a.alloc(msm_reduce_scratch_bytes(b->gz, cap, false)));
    MP_TRY(b->red_b1.alloc(msm_reduce_scratch_bytes(b->gz, cap, false)));
    MP_TRY(b->red_l.alloc(msm_reduce_scratch_bytes(b->gz, cap, false)));
    MP_TRY(b->red_h.alloc(msm_reduce_scratch_bytes(b->gh, cap, false)));
    MP_TRY(b->red_b2.alloc(msm_reduce_scratch_bytes(b->gz, cap, true)));
    MP_TRY(b->proofs.alloc(cap * MP_PROOF_BYTES));
    MP_TRY(b->bad_dev.alloc(4));
    MP_CUDA_TRY(cudaHostAlloc((void**)&b->bad_host, 4, cudaHostAllocDefault));
    *b->bad_host = 0;
    b->device_bytes = dev_alloc_counter() - alloc0;   // every per-proof buffer is a DevBuf allocated above
    return MP_OK;
}

// The view of one MSM job for the proofs [s0, ...) of the batch: everything that is indexed by the proof moves by s0, the point
// buffer of the bucket trees (sized for one slab), the window table and the reduction scratch stay.
static void shift_sort_ws(MsmSortWs& w, const MsmGeom& g, size_t s0) {
    w.cnt += s0 * g.n_buckets;
    w.start += s0 * g.n_buckets;
    w.fill += s0 * g.n_buckets;
    w.slot_base += s0 * (g.n_buckets + 1);
    w.items += s0 * (size_t)g.max_items * 2;
    w.n_items += s0;
    w.entries += s0 * (size_t)g.ent_cap;
    w.heavy += s0 * g.max_heavy;
    w.n_heavy += s0;
    w.q += s0 * (size_t)(g.ba_rounds + 1) * (PLAN_THREADS + 1);
}
// The same job over the vectors [s0, ...) of its sorted list only (everything else - buffers, results - stays).
static MsmJob job_list_view(const MsmJob& j, size_t s0) {
    MsmJob o = j;
    shift_sort_ws(o.ws, j.g, s0);
    return o;
}
static MsmJob job_slab(const MsmJob& j, size_t s0, size_t point_bytes, size_t xyzz_bytes) {
    MsmJob o = j;
    if (s0 == 0) return o;
    shift_sort_ws(o.ws, j.g, s0);
    shift_sort_ws(o.ws_rc, j.g_rc, s0);
    o.result = (char*)j.result + s0 * j.g.groups * xyzz_bytes;
    o.pbuf_rc = (char*)j.pbuf_rc + s0 * (size_t)j.g_rc.p_cap * point_bytes;
    o.result_rc = (char*)j.result_rc + s0 * j.g_rc.groups * xyzz_bytes;
    return o;
}

// Small batches (<= 16 proofs with separate buffers) leave the GPU mostly idle and are bound by the serial steps of three chains,
// each on its own stream:
//   second stream: B list -> G2 bucket trees -> G2 reduction -> G2 bytes
//   third stream : A and L lists (-> for one or two proofs: bucket trees and reduction of A, B1, L [+ the s g_a and r g1_b MSMs])
//   main stream  : witness map -> H list -> H trees -> H reduction -> assembly of g_c
// Everything that needs no read-back (the sorts, the witness map) is enqueued FIRST: the host blocks in msm_ba_rounds_needed
// (one or two proofs: how many tree levels are populated) and the chains behind those read-backs start in the order their
// lists become ready.
static int batch_enqueue_small(mp_batch* b, MsmJob* g1, MsmJob* g2, uint64_t launches0) {
    mp_ctx* c = b->ctx;
    const size_t cnt = b->count, cap = b->capacity;
    cudaStream_t st = b->st, sg2 = b->st2, s3 = b->st3;
    const size_t g1w = XYZZ<Fq>::WORDS * 4;
    const MsmBaWs* ba1 = b->use_ba ? &b->ba_g1 : nullptr;
    const MsmBaWs* ba2 = b->use_ba ? &b->ba_g2 : nullptr;
    const bool trim_rounds = b->use_ba && cnt <= 2;
    const bool split_abl = trim_rounds && b->ba_mem_h.p;
    // MP_LADDERS_AS_MSM=1: s g_a and r g1_b as two more MSM jobs (valid when every A / B1 table point has order r - the GLV
    // condition).  Measured on the B200 (profiles/r02q_*): 4.89 ms per proof against 4.74 with the ladders - the G2 chain is
    // the critical path of a single proof, and the extra MSM work slows it down more than the shorter A / B1 / L chain helps.
    bool ext = false;
    if (const char* e = getenv("MP_LADDERS_AS_MSM")) ext = e[0] == '1' && split_abl && c->glv && b->zx.p;
    const size_t row = (size_t)c->zlen * 8;   // words per scalar vector
    const uint32_t* zc = b->z_canon.as<uint32_t>();
    const uint32_t* zx = b->zx.as<uint32_t>();
    b->ba_g1.round_limit = b->ba_g2.round_limit = b->ba_h.round_limit = b->ba_abl.round_limit = 0;
    MP_CUDA_TRY(cudaEventRecord(b->ev[PH_G2], st));
    if (ext) {
        k_prove_scale<<<dim3(div_up(c->zlen, 256), (unsigned)cnt), 256, 0, st>>>(zc, b->rs.as<uint32_t>(), c->zlen, (uint32_t)cnt, b->zx.as<uint32_t>());
        MP_KERNEL_CHECK();
    }
    MP_CUDA_TRY(cudaEventRecord(b->ev_tail_fork, st));  // z' complete
    MP_CUDA_TRY(cudaStreamWaitEvent(sg2, b->ev_tail_fork, 0));
    MP_CUDA_TRY(cudaStreamWaitEvent(s3, b->ev_tail_fork, 0));
    // ---- lists.  ext: the B list is sorted as 2 cnt vectors (r z' | z'), the A list as (z' | s z')
    nvtxRangePushA("Compute B in G2");
    if (ext) MP_TRY(msm_sort(b->gz, zx, row, 2 * cnt, b->sort_b, c->valid_b.as<uint32_t>(), sg2));
    else MP_TRY(msm_sort(b->gz, zc, row, cnt, b->sort_b, c->valid_b.as<uint32_t>(), sg2));
    MP_CUDA_TRY(cudaEventRecord(b->ev_sort_b, sg2));  // the B1 job of the G1 launch reads the B list
    if (ext) MP_TRY(msm_sort(b->gz, zx + cnt * row, row, 2 * cnt, b->sort_a, c->valid_a.as<uint32_t>(), s3));
    else MP_TRY(msm_sort(b->gz, zc, row, cnt, b->sort_a, c->valid_a.as<uint32_t>(), s3));
    MP_TRY(msm_sort(b->gz, zc, row, cnt, b->sort_l, c->valid_l.as<uint32_t>(), s3));
    MP_CUDA_TRY(cudaEventRecord(b->ev_sort_al, s3));
    nvtxRangePop();
    // ---- witness map and H list
    MP_CUDA_TRY(cudaEventRecord(b->ev[PH_WITNESS_MAP], st));
    nvtxRangePushA("R1CS to QAP witness map");
    if (!b->abc_supplied) MP_TRY(r1cs_eval(c->r1cs, b->z_mont.p, c->zlen, cnt, c->m, b->p_abc, st));
    MP_TRY(witness_map_run(c->dom, b->p_abc, b->p_s1, b->p_s2, cnt, b->h_canon.p, c->m, st));
    nvtxRangePop();
    MP_CUDA_TRY(cudaEventRecord(b->ev[PH_SORT], st));
    MP_TRY(msm_sort(b->gh, b->h_canon.as<uint32_t>(), (size_t)c->m * 8, cnt, b->sort_h, c->valid_h.as<uint32_t>(), st));
    // ---- G2 chain
    {
        nvtxRangePushA("Compute B in G2");
        MsmJob jg2 = ext ? job_list_view(g2[0], cnt) : g2[0];
        if (trim_rounds) MP_TRY(msm_ba_rounds_needed(&jg2, 1, cnt, sg2, &b->ba_g2.round_limit));
        MP_TRY(msm_accumulate_g2(&jg2, 1, cnt, ba2, sg2));
        MP_TRY(msm_reduce_heavy_g2(&jg2, 1, cnt, ba2, sg2));
        MP_TRY(msm_reduce_tail_g2(&jg2, 1, cnt, ba2, sg2));
        k_prove_finish_g2<<<div_up(cnt, 32), 32, 0, sg2>>>(b->res_g2.as<XYZZ<Fq2>>(), (uint32_t)cnt, b->proofs.as<uint8_t>());
        MP_KERNEL_CHECK();
        MP_CUDA_TRY(cudaEventRecord(b->ev_g2, sg2));
        nvtxRangePop();
    }
    // ---- A, B1, L chain of one or two proofs
    nvtxRangePushA("Compute A, Compute B in G1, Compute C (H and L queries)");
    if (split_abl) {
        MP_CUDA_TRY(cudaStreamWaitEvent(s3, b->ev_sort_b, 0));
        MsmJob abl[5];
        int n_abl = 3;
        MsmBaWs* w = &b->ba_g1;
        abl[0] = g1[0];
        abl[1] = ext ? job_list_view(g1[1], cnt) : g1[1];
        abl[2] = g1[2];
        if (ext) {
            char* res1 = b->res_g1.as<char>();
            abl[3] = job_list_view(g1[0], cnt);   // s z' over the A table
            abl[3].pbuf = b->pb_a.as<char>() + cap * b->gz.p_cap * MP_G1_BYTES;
            abl[3].result = res1 + 4 * cnt * g1w;
            abl[3].scratch = b->red_sa.p;
            abl[4] = g1[1];                       // r z' over the B1 table: the first half of the B list
            abl[4].pbuf = b->pb_b1.as<char>() + cap * b->gz.p_cap * MP_G1_BYTES;
            abl[4].result = res1 + 5 * cnt * g1w;
            abl[4].scratch = b->red_rb.p;
            n_abl = 5;
            w = &b->ba_abl;
        }
        MP_TRY(msm_ba_rounds_needed(abl, n_abl, cnt, s3, &w->round_limit));
        MP_TRY(msm_accumulate_g1(abl, n_abl, cnt, w, s3));
        MP_TRY(msm_reduce_heavy_g1(abl, n_abl, cnt, w, s3));
        MP_TRY(msm_reduce_tail_g1(abl, n_abl, cnt, w, s3));
        if (!ext) {
            k_prove_ladders<<<(unsigned)cnt, 96, 0, s3>>>(b->res_g1.as<XYZZ<Fq>>(), b->rs.as<uint32_t>(), (uint32_t)cnt, c->glv ? 1 : 0,
                                                         b->lad.as<uint32_t>(), b->proofs.as<uint8_t>());
            MP_KERNEL_CHECK();
        }
        MP_CUDA_TRY(cudaEventRecord(b->ev_acc_abl, s3));
    } else {
        MP_CUDA_TRY(cudaStreamWaitEvent(st, b->ev_sort_b, 0));
        MP_CUDA_TRY(cudaStreamWaitEvent(st, b->ev_sort_al, 0));
    }
    // ---- H chain (or all four G1 MSMs) on the main stream
    MP_CUDA_TRY(cudaEventRecord(b->ev[PH_ACC_G1], st));
    if (split_abl) {
        MP_TRY(msm_ba_rounds_needed(g1 + 3, 1, cnt, st, &b->ba_h.round_limit));
        MP_TRY(msm_accumulate_g1(g1 + 3, 1, cnt, &b->ba_h, st));
    } else {
        if (trim_rounds) MP_TRY(msm_ba_rounds_needed(g1, 4, cnt, st, &b->ba_g1.round_limit));
        MP_TRY(msm_accumulate_g1(g1, 4, cnt, ba1, st));
    }
    MP_CUDA_TRY(cudaEventRecord(b->ev[PH_REDUCE], st));
    if (split_abl) MP_TRY(msm_reduce_heavy_g1(g1 + 3, 1, cnt, &b->ba_h, st));
    else MP_TRY(msm_reduce_heavy_g1(g1, 4, cnt, ba1, st));
    MP_CUDA_TRY(cudaEventRecord(b->ev_heavy, st));  // the next batch of this context may start its kernels now
    c->last_heavy = b->ev_heavy;
    c->last_heavy_owner = b;
    if (split_abl) MP_TRY(msm_reduce_tail_g1(g1 + 3, 1, cnt, &b->ba_h, st));
    else MP_TRY(msm_reduce_tail_g1(g1, 4, cnt, ba1, st));
    nvtxRangePop();
    MP_CUDA_TRY(cudaEventRecord(b->ev[PH_FINISH], st));
    NvtxRange fin("Finish C");
    if (split_abl) {
        MP_CUDA_TRY(cudaStreamWaitEvent(st, b->ev_acc_abl, 0));
        if (ext) k_prove_assemble_msm<<<(unsigned)cnt, 64, 0, st>>>(b->res_g1.as<XYZZ<Fq>>(), (uint32_t)cnt, b->proofs.as<uint8_t>());
        else k_prove_assemble<<<div_up(cnt, 32), 32, 0, st>>>(b->res_g1.as<XYZZ<Fq>>(), b->lad.as<uint32_t>(), (uint32_t)cnt, b->proofs.as<uint8_t>());
    } else {
        k_prove_finish<<<(unsigned)cnt, FINISH_THREADS, 0, st>>>(b->res_g1.as<XYZZ<Fq>>(), b->rs.as<uint32_t>(), (uint32_t)cnt, c->glv ? 1 : 0,
                                                                 b->proofs.as<uint8_t>());
    }
    MP_KERNEL_CHECK();
    MP_CUDA_TRY(cudaStreamWaitEvent(st, b->ev_g2, 0));   // the G2 bytes of the proofs (k_prove_finish_g2 on the second stream)
    MP_CUDA_TRY(cudaEventRecord(b->ev[PH_COUNT], st));
    b->launches = kernel_launch_counter() - launches0;
    b->in_flight = true;
    return MP_OK;
}

// Enqueues every kernel of one batch on the batch's streams; returns without synchronising.
// Main stream: prep -> G2 MSM (B list) -> R1CS + witness map -> A/L/H lists -> G1 MSMs -> finish.  The latency-bound
// tail of the G2 reduction runs on the second stream beside the witness map and the G1 MSMs.
static int batch_enqueue(mp_batch* b) {
    mp_ctx* c = b->ctx;
    const size_t cnt = b->count;
    if (cnt == 0) return MP_OK;
    MP_TRY(use_device(c->device));
    NvtxRange whole("Groth16::Prover");
    cudaStream_t st = b->st;
    const uint64_t launches0 = kernel_launch_counter();
    const size_t g1w = XYZZ<Fq>::WORDS * 4, g2w = XYZZ<Fq2>::WORDS * 4;
    cudaStream_t st_tail = b->overlap ? b->st2 : st;  // stream of the G2 reduction tail
    if (c->last_heavy && c->last_heavy_owner != b) MP_CUDA_TRY(cudaStreamWaitEvent(st, c->last_heavy, 0));
    MP_CUDA_TRY(cudaEventRecord(b->ev[PH_PREP], st));
    MP_CUDA_TRY(cudaMemsetAsync(b->bad_dev.p, 0, 4, st));
    k_prove_prep<<<dim3(div_up(c->n + 1, 256), (unsigned)cnt), 256, 0, st>>>(b->z_canon.as<uint32_t>(), b->z_mont.as<uint32_t>(),
                                                                           b->rs.as<uint32_t>(), (uint32_t)c->n, c->zlen, b->bad_dev.as<uint32_t>());
    MP_KERNEL_CHECK();
    MP_CUDA_TRY(cudaMemcpyAsync(b->bad_host, b->bad_dev.p, 4, cudaMemcpyDeviceToHost, st));
    char* res1 = b->res_g1.as<char>();
    char* rrc1 = b->resrc_g1.as<char>();
    MsmJob g1[4] = {
        {b->gz, b->sort_a, c->tab_a.p, b->part_a.p, res1, b->red_a.p, b->pb_a.p, b->gz_rc, b->rc_a, b->pbrc_a.p, rrc1},
        {b->gz, b->sort_b, c->tab_b1.p, b->part_b1.p, res1 + cnt * g1w, b->red_b1.p, b->pb_b1.p, b->gz_rc, b->rc_b, b->pbrc_b1.p, rrc1 + 2 * cnt * g1w},
        {b->gz, b->sort_l, c->tab_l.p, b->part_l.p, res1 + 2 * cnt * g1w, b->red_l.p, b->pb_l.p, b->gz_rc, b->rc_l, b->pbrc_l.p, rrc1 + 4 * cnt * g1w},
        {b->gh, b->sort_h, c->tab_h.p, b->part_h.p, res1 + 3 * cnt * g1w, b->red_h.p, b->p_pb_h, b->gh_rc, b->rc_h, b->pbrc_h.p, rrc1 + 6 * cnt * g1w},
    };
    MsmJob g2[1] = {{b->gz, b->sort_b, c->tab_b2.p, b->part_b2.p, b->res_g2.p, b->red_b2.p, b->p_pb_b2, b->gz_rc, b->rc_b2, b->pbrc_b2.p, b->resrc_g2.p}};
    const MsmBaWs* ba1 = b->use_ba ? &b->ba_g1 : nullptr;
    const MsmBaWs* ba2 = b->use_ba ? &b->ba_g2 : nullptr;
    // ---- G2 MSM over the B list.  Large batches: on the main stream (co-running throughput kernels costs ~5 %), only the
    // latency-bound tail of its reduction on the second stream.  Small batches leave the GPU mostly idle and are bound by the
    // per-level latencies of the trees: the whole G2 MSM then runs on the second stream beside the witness map and the G1 MSMs.
    const bool g2_side = b->overlap && cnt <= 16 && !b->aliased;
    if (g2_side) {
        const char* e = getenv("MP_SMALL_PATH_OLD");   // A/B hook: the first form of the small-batch path, below
        if (!(e && e[0] == '1')) return batch_enqueue_small(b, g1, g2, launches0);
    }
    cudaStream_t sg2 = g2_side ? b->st2 : st;
    MP_CUDA_TRY(cudaEventRecord(b->ev[PH_G2], st));
    nvtxRangePushA("Compute B in G2");
    if (g2_side) {
        MP_CUDA_TRY(cudaEventRecord(b->ev_tail_fork, st));  // z' complete
        MP_CUDA_TRY(cudaStreamWaitEvent(sg2, b->ev_tail_fork, 0));
    }
    MP_TRY(msm_sort(b->gz, b->z_canon.as<uint32_t>(), (size_t)c->zlen * 8, cnt, b->sort_b, c->valid_b.as<uint32_t>(), sg2));
    if (g2_side) MP_CUDA_TRY(cudaEventRecord(b->ev_sort_b, sg2));  // the B1 job of the G1 launch reads the B list
    // one or two proofs: find out how many tree levels are populated instead of walking all the provisioned ones
    const bool trim_rounds = b->use_ba && cnt <= 2;
    b->ba_g1.round_limit = b->ba_g2.round_limit = 0;
    if (trim_rounds) MP_TRY(msm_ba_rounds_needed(g2, 1, cnt, sg2, &b->ba_g2.round_limit));
    const size_t slab = b->acc_slabs > 1 ? b->slab_cap : cnt;   // proofs per pass over the point buffers
    for (size_t s0 = 0; s0 < cnt; s0 += slab) {
        const MsmJob js = job_slab(g2[0], s0, MP_G2_BYTES, g2w);
        const size_t n = std::min(slab, cnt - s0);
        MP_TRY(msm_accumulate_g2(&js, 1, n, ba2, sg2));
        MP_TRY(msm_reduce_heavy_g2(&js, 1, n, ba2, sg2));
    }
    if (b->overlap && !g2_side) {
        MP_CUDA_TRY(cudaEventRecord(b->ev_g2_heavy, st));
        MP_CUDA_TRY(cudaStreamWaitEvent(st_tail, b->ev_g2_heavy, 0));
    }
    MP_TRY(msm_reduce_tail_g2(g2, 1, cnt, ba2, st_tail));
    k_prove_finish_g2<<<div_up(cnt, 32), 32, 0, st_tail>>>(b->res_g2.as<XYZZ<Fq2>>(), (uint32_t)cnt, b->proofs.as<uint8_t>());
    MP_KERNEL_CHECK();
    if (b->overlap) MP_CUDA_TRY(cudaEventRecord(b->ev_g2, st_tail));
    nvtxRangePop();
    // a small batch is latency-bound: the A and L lists only need z', so their sorts run beside the witness map
    if (g2_side) {
        MP_CUDA_TRY(cudaStreamWaitEvent(b->st3, b->ev_tail_fork, 0));
        MP_TRY(msm_sort(b->gz, b->z_canon.as<uint32_t>(), (size_t)c->zlen * 8, cnt, b->sort_a, c->valid_a.as<uint32_t>(), b->st3));
        MP_TRY(msm_sort(b->gz, b->z_canon.as<uint32_t>(), (size_t)c->zlen * 8, cnt, b->sort_l, c->valid_l.as<uint32_t>(), b->st3));
        MP_CUDA_TRY(cudaEventRecord(b->ev_sort_al, b->st3));
    }
    // ---- witness map
    MP_CUDA_TRY(cudaEventRecord(b->ev[PH_WITNESS_MAP], st));
    nvtxRangePushA("R1CS to QAP witness map");
    if (!b->abc_supplied) MP_TRY(r1cs_eval(c->r1cs, b->z_mont.p, c->zlen, cnt, c->m, b->p_abc, st));
    MP_TRY(witness_map_run(c->dom, b->p_abc, b->p_s1, b->p_s2, cnt, b->h_canon.p, c->m, st));
    nvtxRangePop();
    // ---- G1 MSMs
    MP_CUDA_TRY(cudaEventRecord(b->ev[PH_SORT], st));
    nvtxRangePushA("Compute A, Compute B in G1, Compute C (H and L queries)");
    if (!g2_side) {
        MP_TRY(msm_sort(b->gz, b->z_canon.as<uint32_t>(), (size_t)c->zlen * 8, cnt, b->sort_a, c->valid_a.as<uint32_t>(), st));
        MP_TRY(msm_sort(b->gz, b->z_canon.as<uint32_t>(), (size_t)c->zlen * 8, cnt, b->sort_l, c->valid_l.as<uint32_t>(), st));
    }
    MP_TRY(msm_sort(b->gh, b->h_canon.as<uint32_t>(), (size_t)c->m * 8, cnt, b->sort_h, c->valid_h.as<uint32_t>(), st));
    if (g2_side) {
        MP_CUDA_TRY(cudaStreamWaitEvent(st, b->ev_sort_b, 0));
        MP_CUDA_TRY(cudaStreamWaitEvent(st, b->ev_sort_al, 0));
    }
    MP_CUDA_TRY(cudaEventRecord(b->ev[PH_ACC_G1], st));
    const bool split_abl = g2_side && trim_rounds && b->ba_mem_h.p;
    if (split_abl) {
        // A, B1, L on the third stream (their lists are ready while the witness map still runs): bucket trees, reduction and the
        // two ladders of the finishing step; H on the main stream; they meet in k_prove_assemble
        MP_CUDA_TRY(cudaStreamWaitEvent(b->st3, b->ev_sort_b, 0));
        MP_TRY(msm_ba_rounds_needed(g1, 3, cnt, b->st3, &b->ba_g1.round_limit));
        MP_TRY(msm_accumulate_g1(g1, 3, cnt, ba1, b->st3));
        MP_TRY(msm_reduce_heavy_g1(g1, 3, cnt, ba1, b->st3));
        MP_TRY(msm_reduce_tail_g1(g1, 3, cnt, ba1, b->st3));
        k_prove_ladders<<<(unsigned)cnt, 96, 0, b->st3>>>(b->res_g1.as<XYZZ<Fq>>(), b->rs.as<uint32_t>(), (uint32_t)cnt, c->glv ? 1 : 0,
                                                         b->lad.as<uint32_t>(), b->proofs.as<uint8_t>());
        MP_KERNEL_CHECK();
        MP_CUDA_TRY(cudaEventRecord(b->ev_acc_abl, b->st3));
        b->ba_h.round_limit = 0;
        MP_TRY(msm_ba_rounds_needed(g1 + 3, 1, cnt, st, &b->ba_h.round_limit));
        MP_TRY(msm_accumulate_g1(g1 + 3, 1, cnt, &b->ba_h, st));
    } else {
        if (trim_rounds) MP_TRY(msm_ba_rounds_needed(g1, 4, cnt, st, &b->ba_g1.round_limit));
        if (b->acc_slabs == 1) MP_TRY(msm_accumulate_g1(g1, 4, cnt, ba1, st));
    }
    MP_CUDA_TRY(cudaEventRecord(b->ev[PH_REDUCE], st));
    if (b->acc_slabs > 1) {   // bucket trees and row/column trees slab after slab (the phase split of the timers ends here)
        for (size_t s0 = 0; s0 < cnt; s0 += slab) {
            MsmJob js[4];
            for (int i = 0; i < 4; i++) js[i] = job_slab(g1[i], s0, MP_G1_BYTES, g1w);
            const size_t n = std::min(slab, cnt - s0);
            MP_TRY(msm_accumulate_g1(js, 4, n, ba1, st));
            MP_TRY(msm_reduce_heavy_g1(js, 4, n, ba1, st));
        }
    } else if (split_abl) {
        MP_TRY(msm_reduce_heavy_g1(g1 + 3, 1, cnt, &b->ba_h, st));
    } else {
        MP_TRY(msm_reduce_heavy_g1(g1, 4, cnt, ba1, st));
    }
    MP_CUDA_TRY(cudaEventRecord(b->ev_heavy, st));  // the next batch of this context may start its kernels now
    c->last_heavy = b->ev_heavy;
    c->last_heavy_owner = b;
    if (split_abl) MP_TRY(msm_reduce_tail_g1(g1 + 3, 1, cnt, &b->ba_h, st));
    else MP_TRY(msm_reduce_tail_g1(g1, 4, cnt, ba1, st));
    nvtxRangePop();
    MP_CUDA_TRY(cudaEventRecord(b->ev[PH_FINISH], st));
    NvtxRange fin("Finish C");
    if (split_abl) {
        MP_CUDA_TRY(cudaStreamWaitEvent(st, b->ev_acc_abl, 0));
        k_prove_assemble<<<div_up(cnt, 32), 32, 0, st>>>(b->res_g1.as<XYZZ<Fq>>(), b->lad.as<uint32_t>(), (uint32_t)cnt, b->proofs.as<uint8_t>());
    } else {
        k_prove_finish<<<(unsigned)cnt, FINISH_THREADS, 0, st>>>(b->res_g1.as<XYZZ<Fq>>(), b->rs.as<uint32_t>(), (uint32_t)cnt, c->glv ? 1 : 0,
                                                                 b->proofs.as<uint8_t>());
    }
    MP_KERNEL_CHECK();
    if (b->overlap) MP_CUDA_TRY(cudaStreamWaitEvent(st, b->ev_g2, 0));   // the G2 bytes of the proofs (k_prove_finish_g2 on the other stream)
    MP_CUDA_TRY(cudaEventRecord(b->ev[PH_COUNT], st));
    b->launches = kernel_launch_counter() - launches0;
    b->in_flight = true;
    return MP_OK;
}

// Waits for the batch's streams and collects the CUDA-event times of the run.
static int batch_finalize(mp_batch* b, float* out_ms) {
    if (out_ms) *out_ms = 0;
    if (!b->in_flight) return MP_OK;
    MP_TRY(use_device(b->ctx->device));
    MP_CUDA_TRY(cudaStreamSynchronize(b->st));
    MP_CUDA_TRY(cudaStreamSynchronize(b->st2));
    MP_CUDA_TRY(cudaStreamSynchronize(b->st3));
    b->in_flight = false;
    float total = 0;
    for (int ph = PH_PREP; ph < PH_COUNT; ph++) MP_CUDA_TRY(cudaEventElapsedTime(&b->phase_ms[ph], b->ev[ph], b->ev[ph + 1]));
    MP_CUDA_TRY(cudaEventElapsedTime(&total, b->ev[PH_PREP], b->ev[PH_COUNT]));
    b->ran = true;
    if (out_ms) *out_ms = total;
    if (*b->bad_host) {
        set_error_detail("proof %u of the batch carries a non-canonical scalar (z, r or s >= the Fr modulus)", *b->bad_host - 1);
        b->ran = false;
        return MP_ERR_INVALID_ARG;
    }
    return MP_OK;
}

}  // namespace mp

// ---------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------
extern "C" {

int mp_pk_parse(const uint8_t* data, size_t len, mp_pk_view* out) {
    if (!data || !out) return MP_ERR_INVALID_ARG;
    size_t pos = 0;
    auto take = [&](size_t nbytes, const uint8_t** p) -> bool {
        if (nbytes > len - pos) return false;
        *p = data + pos;
        pos += nbytes;
        return true;
    };
    auto vec = [&](size_t elem, const uint8_t** p, uint64_t* cnt) -> bool {
        const uint8_t* lp;
        if (!take(8, &lp)) return false;
        uint64_t c = 0;
        memcpy(&c, lp, 8);
        if (c > (len - pos) / elem) return false;
        *cnt = c;
        return take((size_t)c * elem, p);
    };
    memset(out, 0, sizeof(*out));
    bool ok = take(MP_G1_BYTES, &out->alpha_g1) && take(MP_G2_BYTES, &out->beta_g2) && take(MP_G2_BYTES, &out->gamma_g2) &&
              take(MP_G2_BYTES, &out->delta_g2) && vec(MP_G1_BYTES, &out->gamma_abc_g1, &out->gamma_abc_len) &&
              take(MP_G1_BYTES, &out->beta_g1) && take(MP_G1_BYTES, &out->delta_g1) && vec(MP_G1_BYTES, &out->a_query, &out->a_len) &&
              vec(MP_G1_BYTES, &out->b_g1_query, &out->b_g1_len) && vec(MP_G2_BYTES, &out->b_g2_query, &out->b_g2_len) &&
              vec(MP_G1_BYTES, &out->h_query, &out->h_len) && vec(MP_G1_BYTES, &out->l_query, &out->l_len);
    if (!ok || pos != len) {
        mp::set_error_detail("proving key: truncated or trailing bytes (consumed %zu of %zu)", pos, len);
        return MP_ERR_FORMAT;
    }
    return MP_OK;
}

int mp_ctx_create(const mp_pk_view* pk, const mp_r1cs_view* r1cs, int device, mp_ctx** out) {
    if (!pk || !r1cs || !out) return MP_ERR_INVALID_ARG;
    if (!pk->alpha_g1 || !pk->beta_g1 || !pk->beta_g2 || !pk->delta_g1 || !pk->delta_g2 || !pk->a_query || !pk->b_g1_query ||
        !pk->b_g2_query || !pk->h_query || (!pk->l_query && pk->l_len))
        return MP_ERR_INVALID_ARG;
    mp_ctx* c = new (std::nothrow) mp_ctx();
    if (!c) return MP_ERR_OOM;
    int rc = ctx_create_impl(pk, r1cs, device, c);
    if (rc != MP_OK) {
        delete c;
        return rc;
    }
    c->device_bytes += c->dom.n * 32 * 6;
    *out = c;
    return MP_OK;
}

void mp_ctx_destroy(mp_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->single) mp_batch_destroy(ctx->single);
    delete ctx;
}

int mp_ctx_info(const mp_ctx* ctx, uint64_t* n_vars, uint64_t* n_instance, uint64_t* domain_size, uint64_t* device_bytes) {
    if (!ctx) return MP_ERR_INVALID_ARG;
    if (n_vars) *n_vars = ctx->n;
    if (n_instance) *n_instance = ctx->p;
    if (domain_size) *domain_size = ctx->m;
    if (device_bytes) *device_bytes = ctx->device_bytes;
    return MP_OK;
}

int mp_batch_create(mp_ctx* ctx, size_t capacity, mp_batch** out) { return mp_batch_create_ex(ctx, capacity, 0, out); }

int mp_batch_create_ex(mp_ctx* ctx, size_t capacity, int high_priority, mp_batch** out) {
    if (!ctx || !out || capacity == 0) return MP_ERR_INVALID_ARG;
    if (capacity > MP_MAX_BATCH) {  // the NTT launches carry 3 * count vectors in gridDim.y (<= 65535)
        mp::set_error_detail("batch capacity %zu exceeds %d", capacity, MP_MAX_BATCH);
        return MP_ERR_UNSUPPORTED;
    }
    size_t slabs = 1;
    if (const char* e = getenv("MP_ACC_SLABS")) {
        const long v = atol(e);
        slabs = v >= 4 ? 4 : (v >= 2 ? 2 : 1);
    }
    for (;; slabs *= 2) {   // out of memory: halve the point buffers (two, then four passes over them) before giving up
        mp_batch* b = new (std::nothrow) mp_batch();
        if (!b) return MP_ERR_OOM;
        int rc = batch_create_impl(ctx, capacity, high_priority, slabs, b);
        if (rc == MP_OK) {
            *out = b;
            return MP_OK;
        }
        mp_batch_destroy(b);
        if (rc != MP_ERR_OOM || slabs >= 4 || capacity <= 16) return rc;
        cudaGetLastError();  // clear the sticky allocation error
    }
}

void mp_batch_destroy(mp_batch* b) {
    if (!b) return;
    if (b->ctx) cudaSetDevice(b->ctx->device);
    if (b->st) cudaStreamSynchronize(b->st);
    if (b->st2) cudaStreamSynchronize(b->st2);
    if (b->st3) cudaStreamSynchronize(b->st3);
    if (b->ev_sort_al) cudaEventDestroy(b->ev_sort_al);
    if (b->ev_acc_abl) cudaEventDestroy(b->ev_acc_abl);
    for (auto& e : b->ev)
        if (e) cudaEventDestroy(e);
    if (b->ctx) {
        std::lock_guard<std::mutex> lock(b->ctx->mu);
        if (b->ctx->last_heavy_owner == b) {
            b->ctx->last_heavy = nullptr;
            b->ctx->last_heavy_owner = nullptr;
        }
    }
    if (b->bad_host) cudaFreeHost(b->bad_host);
    for (cudaEvent_t e : {b->ev_g2_heavy, b->ev_g2, b->ev_heavy, b->ev_tail_fork, b->ev_sort_b, b->ev_dom0, b->ev_dom1})
        if (e) cudaEventDestroy(e);
    for (auto e : b->ev_pipe)
        if (e) cudaEventDestroy(e);
    for (auto st : b->st_mid)
        if (st) cudaStreamDestroy(st);
    if (b->st) cudaStreamDestroy(b->st);
    if (b->st2) cudaStreamDestroy(b->st2);
    if (b->st3) cudaStreamDestroy(b->st3);
    delete b;
}

int mp_batch_upload(mp_batch* b, size_t count, const uint64_t* z, const uint64_t* r, const uint64_t* s) {
    if (!b || b->in_flight || count > b->capacity || (count && (!z || !r || !s))) return MP_ERR_INVALID_ARG;
    mp_ctx* c = b->ctx;
    MP_TRY(use_device(c->device));
    MP_CUDA_TRY(cudaEventRecord(b->ev[PH_UPLOAD], b->st));
    if (count) {
        MP_CUDA_TRY(cudaMemcpy2DAsync(b->z_canon.p, (size_t)c->zlen * 32, z, c->n * 32, c->n * 32, count, cudaMemcpyHostToDevice, b->st));
        MP_CUDA_TRY(cudaMemcpy2DAsync(b->rs.p, 64, r, 32, 32, count, cudaMemcpyHostToDevice, b->st));
        MP_CUDA_TRY(cudaMemcpy2DAsync(b->rs.as<char>() + 32, 64, s, 32, 32, count, cudaMemcpyHostToDevice, b->st));
    }
    MP_CUDA_TRY(cudaEventRecord(b->ev[PH_PREP], b->st));
    MP_CUDA_TRY(cudaStreamSynchronize(b->st));
    MP_CUDA_TRY(cudaEventElapsedTime(&b->phase_ms[PH_UPLOAD], b->ev[PH_UPLOAD], b->ev[PH_PREP]));
    b->count = count;
    b->ran = false;
    return MP_OK;
}

int mp_batch_run(mp_batch* b, float* out_device_ms) {
    if (!b) return MP_ERR_INVALID_ARG;
    MP_TRY(mp_batch_run_async(b));
    return mp_batch_wait(b, out_device_ms);
}

int mp_batch_run_async(mp_batch* b) {
    if (!b || b->in_flight) return MP_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(b->ctx->mu);
    return batch_enqueue(b);
}

int mp_batch_wait(mp_batch* b, float* out_device_ms) {
    if (!b) return MP_ERR_INVALID_ARG;
    return batch_finalize(b, out_device_ms);
}

int mp_batch_submit(mp_batch* b, size_t count, const uint64_t* z, const uint64_t* r, const uint64_t* s, uint8_t* out_proofs) {
    if (!b || b->in_flight || count > b->capacity || (count && (!z || !r || !s || !out_proofs))) return MP_ERR_INVALID_ARG;
    mp_ctx* c = b->ctx;
    MP_TRY(use_device(c->device));
    std::lock_guard<std::mutex> lock(c->mu);
    b->count = count;
    b->ran = false;
    if (count == 0) return MP_OK;
    MP_CUDA_TRY(cudaMemcpy2DAsync(b->z_canon.p, (size_t)c->zlen * 32, z, c->n * 32, c->n * 32, count, cudaMemcpyHostToDevice, b->st));
    MP_CUDA_TRY(cudaMemcpy2DAsync(b->rs.p, 64, r, 32, 32, count, cudaMemcpyHostToDevice, b->st));
    MP_CUDA_TRY(cudaMemcpy2DAsync(b->rs.as<char>() + 32, 64, s, 32, 32, count, cudaMemcpyHostToDevice, b->st));
    MP_TRY(batch_enqueue(b));
    MP_CUDA_TRY(cudaMemcpyAsync(out_proofs, b->proofs.p, count * MP_PROOF_BYTES, cudaMemcpyDeviceToHost, b->st));
    return MP_OK;
}

int mp_batch_download(mp_batch* b, uint8_t* out_proofs) {
    if (!b || (b->count && !out_proofs)) return MP_ERR_INVALID_ARG;
    if (b->in_flight) MP_TRY(batch_finalize(b, nullptr));
    if (!b->ran && b->count) { mp::set_error_detail("mp_batch_download before mp_batch_run"); return MP_ERR_INVALID_ARG; }
    MP_TRY(use_device(b->ctx->device));
    if (b->count) MP_CUDA_TRY(cudaMemcpy(out_proofs, b->proofs.p, b->count * MP_PROOF_BYTES, cudaMemcpyDeviceToHost));
    return MP_OK;
}

int mp_batch_phase_ms(const mp_batch* b, float* out_ms, int max_phases) {
    if (!b || !out_ms) return 0;
    int nph = max_phases < PH_COUNT ? max_phases : PH_COUNT;
    for (int i = 0; i < nph; i++) out_ms[i] = b->phase_ms[i];
    return nph;
}

int mp_batch_dominant_kernel(mp_batch* b, float* out_ms, uint64_t* out_additions) {
    if (!b || b->in_flight || !b->ran || !b->use_ba) return MP_ERR_INVALID_ARG;
    MP_TRY(mp::use_device(b->ctx->device));
    if (out_ms) MP_CUDA_TRY(cudaEventElapsedTime(out_ms, b->ev_dom0, b->ev_dom1));
    if (out_additions) {
        // pairs of tree level 1 = last entry of row 1 of every list's q table; the B list feeds the B1 job
        uint64_t total = 0;
        const MsmSortWs* lists[4] = {&b->sort_a, &b->sort_b, &b->sort_l, &b->sort_h};
        const MsmGeom* geoms[4] = {&b->gz, &b->gz, &b->gz, &b->gh};
        // the bracketed launch covers the first slab of the batch (msm_impl.inc: a tree level of a large batch runs in slabs)
        const size_t covered = b->ba_g1.dom_count ? std::min(b->ba_g1.dom_count, b->count) : b->count;
        std::vector<uint32_t> host(covered);
        for (int i = 0; i < 4; i++) {
            const size_t row = (size_t)(geoms[i]->ba_rounds + 1) * (PLAN_THREADS + 1);
            MP_CUDA_TRY(cudaMemcpy2D(host.data(), 4, lists[i]->q + (size_t)(PLAN_THREADS + 1) + PLAN_THREADS, row * 4, 4, covered,
                                     cudaMemcpyDeviceToHost));
            for (uint32_t v : host) total += v;
        }
        *out_additions = total;
    }
    return MP_OK;
}

const char* mp_phase_name(int i) { return (i >= 0 && i < PH_COUNT) ? kPhaseNames[i] : ""; }

uint64_t mp_batch_kernel_launches(const mp_batch* b) { return b ? b->launches : 0; }
uint64_t mp_batch_device_bytes(const mp_batch* b) { return b ? b->device_bytes : 0; }

int mp_batch_set_overlap(mp_batch* b, int overlap) {
    if (!b) return MP_ERR_INVALID_ARG;
    b->overlap = overlap != 0;
    return MP_OK;
}

int mp_prove_batch(mp_ctx* ctx, size_t count, const uint64_t* z, const uint64_t* r, const uint64_t* s, uint8_t* out_proofs) {
    if (!ctx || (count && (!z || !r || !s || !out_proofs))) return MP_ERR_INVALID_ARG;
    if (count == 0) return MP_OK;
    // One batch object of at most 128 proofs (~0.4 GB of device buffers per proof) is reused over chunks of the request;
    // when the device is short of memory the chunk is halved until the buffers fit.
    size_t chunk = 128;
    if (const char* e = getenv("MP_PROVE_BATCH_CHUNK")) {  // test hook / memory knob
        long v = atol(e);
        if (v >= 1 && v <= MP_MAX_BATCH) chunk = (size_t)v;
    }
    if (count < chunk) chunk = count;
    mp_batch* b = nullptr;
    int rc;
    while ((rc = mp_batch_create(ctx, chunk, &b)) == MP_ERR_OOM && chunk > 1) {
        cudaGetLastError();  // clear the sticky allocation error
        chunk = (chunk + 1) / 2;
    }
    if (rc != MP_OK) return rc;
    const size_t n = ctx->n;
    for (size_t done = 0; done < count && rc == MP_OK; done += chunk) {
        const size_t c = count - done < chunk ? count - done : chunk;
        rc = mp_batch_upload(b, c, z + done * n * 4, r + done * 4, s + done * 4);
        if (rc == MP_OK) rc = mp_batch_run(b, nullptr);
        if (rc == MP_OK) rc = mp_batch_download(b, out_proofs + done * MP_PROOF_BYTES);
    }
    mp_batch_destroy(b);
    return rc;
}

int mp_prove(mp_ctx* ctx, const uint64_t* z, const uint64_t r[4], const uint64_t s[4], uint8_t out_proof[MP_PROOF_BYTES]) {
    if (!ctx || !z || !r || !s || !out_proof) return MP_ERR_INVALID_ARG;
    // one cached capacity-1 batch per context: no allocation on the per-proof path; concurrent callers queue here
    std::lock_guard<std::mutex> lock(ctx->single_mu);
    if (!ctx->single) MP_TRY(mp_batch_create(ctx, 1, &ctx->single));
    MP_TRY(mp_batch_upload(ctx->single, 1, z, r, s));
    MP_TRY(mp_batch_run(ctx->single, nullptr));
    return mp_batch_download(ctx->single, out_proof);
}

int mp_debug_prove_ba_demand(uint32_t n_vars, uint32_t domain_size, size_t capacity, size_t count, int g2, uint64_t out[5]) {
    if (!out || !n_vars || !domain_size || !capacity || count > capacity) return MP_ERR_INVALID_ARG;
    const uint32_t zlen = n_vars + N_EXTRA;
    MsmGeom gz = msm_geom(PROVE_C, 1, zlen, zlen, capacity * 2), gh = msm_geom(PROVE_C, 1, domain_size, domain_size, capacity * 2);
    const MsmGeom g1[4] = {gz, gz, gz, gh};
    const size_t slab = msm_ba_slab(capacity, g2 != 0);
    if (g2) msm_ba_ws_demand(&gz, 1, slab, std::min(count, slab), out);
    else msm_ba_ws_demand(g1, 4, slab, std::min(count, slab), out);
    out[4] = slab;
    return MP_OK;
}

int mp_prove_from_abc(mp_ctx* ctx, const uint64_t* z, const uint64_t* a, const uint64_t* b, const uint64_t* c, const uint64_t r[4],
                      const uint64_t s[4], uint8_t out_proof[MP_PROOF_BYTES]) {
    if (!ctx || !z || !a || !b || !c || !r || !s || !out_proof) return MP_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->single_mu);
    if (!ctx->single) MP_TRY(mp_batch_create(ctx, 1, &ctx->single));
    mp_batch* bt = ctx->single;
    MP_TRY(mp_batch_upload(bt, 1, z, r, s));
    const size_t vec = ctx->m * 32;
    const uint64_t* src[3] = {a, b, c};
    for (int i = 0; i < 3; i++) MP_CUDA_TRY(cudaMemcpyAsync(bt->abc.as<char>() + i * vec, src[i], vec, cudaMemcpyHostToDevice, bt->st));
    MP_TRY(fr_to_mont(bt->abc.p, bt->abc.p, 3 * ctx->m, bt->st));
    bt->abc_supplied = true;
    int rc = mp_batch_run(bt, nullptr);
    bt->abc_supplied = false;
    MP_TRY(rc);
    return mp_batch_download(bt, out_proof);
}

int mp_witness_map(mp_ctx* ctx, const uint64_t* z, uint64_t* out_h) {
    if (!ctx || !z || !out_h) return MP_ERR_INVALID_ARG;
    MP_TRY(use_device(ctx->device));
    std::lock_guard<std::mutex> lock(ctx->mu);
    const size_t m = ctx->m;
    DevBuf zc, zm, abc, s1, s2, h;
    MP_TRY(zc.alloc(ctx->n * 32));
    MP_TRY(zm.alloc(ctx->n * 32));
    MP_TRY(abc.alloc(3 * m * 32));
    MP_TRY(s1.alloc(3 * m * 32));
    MP_TRY(s2.alloc(3 * m * 32));
    MP_TRY(h.alloc(m * 32));
    MP_CUDA_TRY(cudaMemcpy(zc.p, z, ctx->n * 32, cudaMemcpyHostToDevice));
    MP_TRY(fr_to_mont(zc.p, zm.p, ctx->n, 0));
    MP_TRY(r1cs_eval(ctx->r1cs, zm.p, ctx->n, 1, m, abc.p, 0));
    MP_TRY(witness_map_run(ctx->dom, abc.p, s1.p, s2.p, 1, h.p, m, 0));
    MP_CUDA_TRY(cudaMemcpy(out_h, h.p, m * 32, cudaMemcpyDeviceToHost));
    return MP_OK;
}

}  // extern "C"
