// Radix-2 NTT over the BLS12-381 scalar field and the R1CS -> QAP witness map on sm_100a.
//
// Replaces ark-poly 0.3 `Radix2EvaluationDomain` in-place transforms and ark-groth16 0.3
// `R1CStoQAP::witness_map` (reached from manta-crypto/src/arkworks/groth16.rs:597).
//
// A transform of n = n1 * n2 points is done in two shared-memory passes (the classic four-step split):
//   pass 1: for every column i2, an n1-point transform over i1 (stride n2), then the twiddle w_n^(i2*k1);
//   pass 2: for every row k1, an n2-point transform over i2; the store scatters to natural order k1 + n1*k2.
// Each pass stages its 2048-element tile and the n/2 twiddles of its sub-transform in shared memory, so the
// data moves HBM -> SMEM -> HBM exactly twice per transform (once for n <= 1024).  Coset shifts, the 1/n of the
// inverse, the division by the vanishing polynomial and the final Montgomery -> canonical conversion are all
// folded into per-element "post" tables applied in the store of pass 2.
#include <algorithm>
#include <vector>

#include "ntt.cuh"
#include "tma.cuh"

namespace mp {

constexpr int NTT_TILE = 2048;      // Fr elements per block tile (64 KiB)
constexpr int NTT_THREADS = 256;
#ifndef MP_NTT_BLOCKS
#define MP_NTT_BLOCKS 3   // resident blocks per SM the register allocation aims at (2 measured no better)
#endif
constexpr unsigned NTT_MAX_SUB = 10;  // largest in-SMEM sub-transform (2^10 points)
constexpr unsigned NTT_MAX_LOG = 26;  // two passes up to 2^20, three up to 2^26 (kzg.rs:43-44 powers; SURVEY.md 5 size axis)

struct NttArgs {
    const uint32_t* in;
    uint32_t* out;
    const uint32_t* tw;     // w^k, k < n
    const uint32_t* pre;    // nullable, n entries
    const uint32_t* post;   // nullable, n entries
    const uint32_t* post_c; // nullable, 1 entry
    unsigned log_n, l1, l2; // n = 2^log_n = 2^l1 * 2^l2
    size_t in_stride, out_stride;  // elements between consecutive vectors
    // Nested use (transforms above 2^20 points): this launch runs the 2^outer_l1 row transforms of an outer four-step split.
    // Vector v = V * 2^outer_l1 + k1 is row k1 of outer transform V; its twiddles are w_N^(k << outer_l1) (tw belongs to the
    // OUTER domain of N = n * 2^outer_l1 points) and its output k lands at index k1 + (k << outer_l1) of outer vector V.
    unsigned outer_l1;
};

MP_DEV unsigned brev_bits(unsigned p, unsigned lg) { return lg ? (__brev(p) >> (32 - lg)) : 0u; }

// ---- TMA bulk copies (tma.cuh) ---------------------------------------------------------------------------------------
// A tile of either pass is a set of contiguous global rows (8 KiB rows in pass 2, 256-byte rows in pass 1): each row
// is one bulk copy issued by a lane of warp 0, so the copy engine - not 256 threads doing LDG + STS - stages the tile.

// In-SMEM decimation-in-frequency transform of `seqs` sequences of 2^lg points.
// element (j, c) at s[(j * sj + c * sc) * 8]; twiddle w_np^e at tws[e * 8].  Leaves bit-reversed order.
MP_DEV void smem_dif(uint32_t* s, const uint32_t* tws, unsigned lg, unsigned lseq, unsigned sj, unsigned sc, bool seq_fastest) {
    const unsigned np = 1u << lg, seqs = 1u << lseq;  // sequence counts are powers of two: no integer division in the loops
    // Two levels per shared-memory round trip (radix-4 step: the four elements i0, i0 + q, i0 + 2q, i0 + 3q of a block stay in
    // registers between the levels), one plain radix-2 level at the end when lg is odd.  Same multiplications, half the
    // LDS/STS traffic and barriers.
    unsigned st = 0;
    for (; st + 1 < lg; st += 2) {
        const unsigned lh = lg - 1 - st, half = 1u << lh, quarter = half >> 1;
        const unsigned quads = (np >> 2) * seqs;
        for (unsigned e = threadIdx.x; e < quads; e += NTT_THREADS) {
            unsigned c, q;
            if (seq_fastest) { c = e & (seqs - 1); q = e >> lseq; }
            else { q = e & ((np >> 2) - 1); c = e >> (lg - 2); }
            const unsigned jj = q & (quarter - 1), blk = q >> (lh - 1);
            const unsigned i0 = (blk << (lh + 1)) + jj;
            uint32_t* p0 = s + (size_t)(i0 * sj + c * sc) * 8;
            uint32_t* p1 = s + (size_t)((i0 + quarter) * sj + c * sc) * 8;
            uint32_t* p2 = s + (size_t)((i0 + half) * sj + c * sc) * 8;
            uint32_t* p3 = s + (size_t)((i0 + half + quarter) * sj + c * sc) * 8;
            const Fr a0 = Fr::load(p0), a1 = Fr::load(p1), a2 = Fr::load(p2), a3 = Fr::load(p3);
            // level st: (a0, a2) with w^(jj << st), (a1, a3) with w^((jj + quarter) << st)
            Fr u0 = a0 + a2, d0 = a0 - a2, u1 = a1 + a3, d1 = a1 - a3;
            if (jj) d0 = d0 * Fr::load(tws + (size_t)(jj << st) * 8);
            d1 = d1 * Fr::load(tws + (size_t)((jj + quarter) << st) * 8);
            // level st + 1: (u0, u1) and (d0, d1), both with w^(jj << (st + 1))
            Fr o1 = u0 - u1, o3 = d0 - d1;
            if (jj) {
                const Fr t = Fr::load(tws + (size_t)(jj << (st + 1)) * 8);
                o1 = o1 * t;
                o3 = o3 * t;
            }
            (u0 + u1).store(p0);
            o1.store(p1);
            (d0 + d1).store(p2);
            o3.store(p3);
        }
        __syncthreads();
    }
    const unsigned total = (np >> 1) * seqs;
    for (; st < lg; st++) {
        const unsigned lh = lg - 1 - st, half = 1u << lh;
        for (unsigned e = threadIdx.x; e < total; e += NTT_THREADS) {
            unsigned c, bf;
            if (seq_fastest) { c = e & (seqs - 1); bf = e >> lseq; }
            else { bf = e & ((np >> 1) - 1); c = e >> (lg - 1); }
            unsigned jj = bf & (half - 1), blk = bf >> lh;
            unsigned i0 = (blk << (lh + 1)) + jj, i1 = i0 + half;
            uint32_t* p0 = s + (size_t)(i0 * sj + c * sc) * 8;
            uint32_t* p1 = s + (size_t)(i1 * sj + c * sc) * 8;
            Fr u = Fr::load(p0), v = Fr::load(p1);
            (u + v).store(p0);
            Fr d = u - v;
            if (jj) d = d * Fr::load(tws + (size_t)(jj << st) * 8);
            d.store(p1);
        }
        __syncthreads();
    }
}

// pass 1 (columns): blockIdx.x = column tile, blockIdx.y = vector
__global__ void __launch_bounds__(NTT_THREADS, MP_NTT_BLOCKS) k_ntt_cols(NttArgs a) {
    extern __shared__ __align__(16) uint32_t smem[];
    const unsigned n1 = 1u << a.l1, n2 = 1u << a.l2;
    const unsigned lC = min(11u - a.l1, a.l2), C = 1u << lC;  // = min(NTT_TILE >> l1, n2)
    const unsigned c0 = blockIdx.x * C;
    const size_t v = blockIdx.y;
    uint32_t* tile = smem;
    uint32_t* tws = smem + (size_t)NTT_TILE * 8;
    const uint32_t* in = a.in + v * a.in_stride * 8;
    for (unsigned e = threadIdx.x; e < (n1 >> 1); e += NTT_THREADS) Fr::load(a.tw + (((size_t)e << a.l2) << a.outer_l1) * 8).store(tws + (size_t)e * 8);
    __shared__ __align__(8) uint64_t bar;
    if (!a.pre) {
        // tile[j][c] <- in[j * n2 + c0 + c]: n1 rows of C * 32 bytes, one bulk copy each
        if (threadIdx.x == 0) mbar_init(&bar, 1);
        __syncthreads();
        if (threadIdx.x < 32) {
            if (threadIdx.x == 0) mbar_arrive_expect_tx(&bar, n1 * C * 32);
            __syncwarp();
            for (unsigned j = threadIdx.x; j < n1; j += 32) bulk_copy_g2s(tile + (size_t)j * C * 8, in + ((size_t)j * n2 + c0) * 8, C * 32, &bar);
        }
        mbar_wait(&bar, 0);
    } else {
        for (unsigned idx = threadIdx.x; idx < n1 * C; idx += NTT_THREADS) {
            unsigned c = idx & (C - 1), j = idx >> lC;
            size_t gi = (size_t)j * n2 + c0 + c;
            Fr x = Fr::load(in + gi * 8) * Fr::load(a.pre + gi * 8);
            x.store(tile + (size_t)idx * 8);
        }
    }
    __syncthreads();
    smem_dif(tile, tws, a.l1, lC, C, 1, true);
    uint32_t* out = a.out + v * a.in_stride * 8;
    const unsigned nmask = (1u << a.log_n) - 1;
    for (unsigned idx = threadIdx.x; idx < n1 * C; idx += NTT_THREADS) {
        unsigned c = idx & (C - 1), p = idx >> lC;
        unsigned k1 = brev_bits(p, a.l1);
        Fr x = Fr::load(tile + (size_t)idx * 8);
        unsigned te = (k1 * (c0 + c)) & nmask;
        if (te) x = x * Fr::load(a.tw + ((size_t)te << a.outer_l1) * 8);
        x.store(out + ((size_t)k1 * n2 + c0 + c) * 8);
    }
}

// pass 2 (rows) — also the whole transform when l1 == 0
__global__ void __launch_bounds__(NTT_THREADS, MP_NTT_BLOCKS) k_ntt_rows(NttArgs a) {
    extern __shared__ __align__(16) uint32_t smem[];
    const unsigned n1 = 1u << a.l1, n2 = 1u << a.l2;
    const unsigned lR = min(11u - a.l2, a.l1), R = 1u << lR;  // = min(NTT_TILE >> l2, n1)
    const unsigned r0 = blockIdx.x * R;
    const size_t v = blockIdx.y;
    const unsigned rs = n2 + 1;  // padded row stride (elements)
    uint32_t* tile = smem;
    uint32_t* tws = smem + (size_t)(NTT_TILE + 64) * 8;
    const uint32_t* in = a.in + v * a.in_stride * 8;
    for (unsigned e = threadIdx.x; e < (n2 >> 1); e += NTT_THREADS) Fr::load(a.tw + (((size_t)e << a.l1) << a.outer_l1) * 8).store(tws + (size_t)e * 8);
    __shared__ __align__(8) uint64_t bar;
    if (!(a.l1 == 0 && a.pre)) {
        // tile row c (padded stride rs) <- the n2 contiguous elements of global row r0 + c: one bulk copy per row
        if (threadIdx.x == 0) mbar_init(&bar, 1);
        __syncthreads();
        if (threadIdx.x < 32) {
            if (threadIdx.x == 0) mbar_arrive_expect_tx(&bar, R * n2 * 32);
            __syncwarp();
            for (unsigned c = threadIdx.x; c < R; c += 32) bulk_copy_g2s(tile + (size_t)c * rs * 8, in + (size_t)(r0 + c) * n2 * 8, n2 * 32, &bar);
        }
        mbar_wait(&bar, 0);
    } else {
        for (unsigned idx = threadIdx.x; idx < n2 * R; idx += NTT_THREADS) {
            unsigned j = idx & (n2 - 1), c = idx >> a.l2;
            size_t gi = (size_t)(r0 + c) * n2 + j;
            Fr x = Fr::load(in + gi * 8) * Fr::load(a.pre + gi * 8);
            x.store(tile + (size_t)(c * rs + j) * 8);
        }
    }
    __syncthreads();
    smem_dif(tile, tws, a.l2, lR, 1, rs, false);
    const size_t k1o = v & (((size_t)1 << a.outer_l1) - 1);   // row of the outer split (0 when not nested)
    uint32_t* out = a.out + (v >> a.outer_l1) * a.out_stride * 8;
    Fr pc;
    if (a.post_c) pc = Fr::load(a.post_c);
    for (unsigned idx = threadIdx.x; idx < n2 * R; idx += NTT_THREADS) {
        unsigned c = idx & (R - 1), p = idx >> lR;
        unsigned k2 = brev_bits(p, a.l2);
        const size_t k = k1o + (((size_t)(r0 + c) + ((size_t)k2 << a.l1)) << a.outer_l1);
        Fr x = Fr::load(tile + (size_t)(c * rs + p) * 8);
        if (a.post) x = x * Fr::load(a.post + k * 8);
        else if (a.post_c) x = x * pc;
        x.store(out + k * 8);
    }
}

static size_t ntt_smem_bytes(unsigned lsub) { return ((size_t)NTT_TILE + 64 + (size_t)(1u << lsub) / 2) * 32; }

// The two shared-memory passes of one (possibly nested) transform of 2^log_n points per vector.
static int ntt_two_pass(NttArgs a, unsigned log_n, const void* in, void* out, void* tmp, size_t count, size_t in_stride, size_t out_stride,
                        cudaStream_t st) {
    a.log_n = log_n;
    if (log_n <= NTT_MAX_SUB) {
        a.l1 = 0;
        a.l2 = log_n;
        a.in = (const uint32_t*)in;
        a.out = (uint32_t*)out;
        a.in_stride = in_stride;
        a.out_stride = out_stride;
        for (size_t v0 = 0; v0 < count; v0 += 65535) {   // gridDim.y limit
            NttArgs b = a;
            b.in += v0 * in_stride * 8;
            if (a.outer_l1 == 0) b.out += v0 * out_stride * 8;
            else if (v0) return MP_ERR_UNSUPPORTED;
            k_ntt_rows<<<dim3(1, (unsigned)std::min<size_t>(65535, count - v0)), NTT_THREADS, ntt_smem_bytes(a.l2), st>>>(b);
            MP_KERNEL_CHECK();
        }
        return MP_OK;
    }
    a.l1 = (log_n + 1) / 2;
    a.l2 = log_n - a.l1;
    const unsigned n1 = 1u << a.l1, n2 = 1u << a.l2;
    const unsigned C = std::min<unsigned>(NTT_TILE >> a.l1, n2), R = std::min<unsigned>(NTT_TILE >> a.l2, n1);
    if (count > 65535) return MP_ERR_UNSUPPORTED;
    NttArgs p1 = a;
    p1.in = (const uint32_t*)in;
    p1.out = (uint32_t*)tmp;
    p1.in_stride = in_stride;   // tmp uses the same stride as in
    p1.out_stride = in_stride;
    k_ntt_cols<<<dim3(n2 / C, (unsigned)count), NTT_THREADS, ntt_smem_bytes(a.l1), st>>>(p1);
    MP_KERNEL_CHECK();
    NttArgs p2 = a;
    p2.in = (const uint32_t*)tmp;
    p2.out = (uint32_t*)out;
    p2.in_stride = in_stride;
    p2.out_stride = out_stride;
    p2.pre = nullptr;
    k_ntt_rows<<<dim3(n1 / R, (unsigned)count), NTT_THREADS, ntt_smem_bytes(a.l2), st>>>(p2);
    MP_KERNEL_CHECK();
    return MP_OK;
}

// Up to 2^20 points: two passes.  Above (<= 2^NTT_MAX_LOG): one more column pass in front - an outer four-step split
// N = 2^o x 2^20 whose 2^o row transforms of 2^20 points are the nested two-pass launches; needs the second scratch `tmp2`.
int ntt_run_strided(const NttDomain& d, bool inverse, const void* in, void* out, void* tmp, size_t count, size_t in_stride,
                    size_t out_stride, const void* pre, const void* post, const void* post_c, cudaStream_t st, void* tmp2) {
    if (count == 0) return MP_OK;
    MP_CUDA_TRY(cudaFuncSetAttribute(k_ntt_cols, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ntt_smem_bytes(NTT_MAX_SUB)));
    MP_CUDA_TRY(cudaFuncSetAttribute(k_ntt_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ntt_smem_bytes(NTT_MAX_SUB)));
    NttArgs a{};
    a.tw = (const uint32_t*)(inverse ? d.tw_inv.p : d.tw_fwd.p);
    a.pre = (const uint32_t*)pre;
    a.post = (const uint32_t*)post;
    a.post_c = (const uint32_t*)post_c;
    if (d.log_n <= 2 * NTT_MAX_SUB) return ntt_two_pass(a, d.log_n, in, out, tmp, count, in_stride, out_stride, st);
    if (!tmp2 || in_stride != d.n) { set_error_detail("ntt: 2^%u points need the second scratch buffer and contiguous vectors", d.log_n); return MP_ERR_UNSUPPORTED; }
    const unsigned o = d.log_n - 2 * NTT_MAX_SUB, li = 2 * NTT_MAX_SUB;   // outer rows 2^o (o <= NTT_MAX_SUB), inner 2^20
    if (count << o > 65535) return MP_ERR_UNSUPPORTED;
    // outer pass 1: for every column i2 < 2^li, the 2^o-point transform over stride 2^li, times w_N^(i2 k1); rows k1 stay contiguous
    NttArgs p1 = a;
    p1.log_n = d.log_n;
    p1.l1 = o;
    p1.l2 = li;
    p1.in = (const uint32_t*)in;
    p1.out = (uint32_t*)tmp;
    p1.in_stride = p1.out_stride = in_stride;
    const unsigned C = NTT_TILE >> o;
    k_ntt_cols<<<dim3((1u << li) / C, (unsigned)count), NTT_THREADS, ntt_smem_bytes(o), st>>>(p1);
    MP_KERNEL_CHECK();
    // the 2^o rows of every vector: 2^li-point transforms with the outer domain's twiddles, scattered to k1 + (k << o)
    NttArgs in2 = a;
    in2.pre = nullptr;
    in2.outer_l1 = o;
    return ntt_two_pass(in2, li, tmp, out, tmp2, count << o, (size_t)1 << li, out_stride, st);
}

int ntt_run(const NttDomain& d, bool inverse, const void* in, void* out, void* tmp, size_t count, const void* pre,
            const void* post, const void* post_c, cudaStream_t st) {
    return ntt_run_strided(d, inverse, in, out, tmp, count, d.n, d.n, pre, post, post_c, st, nullptr);
}

// ---------------------------------------------------------------------------------------------------------
// domain tables
// ---------------------------------------------------------------------------------------------------------
MP_COLD Fr fr_pow_u64(Fr base, uint64_t e) {
    Fr r = Fr::one();
    while (e) {
        if (e & 1) r = r * base;
        base = base.sqr();
        e >>= 1;
    }
    return r;
}

// tw_fwd[k] = w^k, tw_inv[k] = w^-k, and the post tables (see NttDomain)
__global__ void k_ntt_tables(unsigned log_n, uint32_t* tw_fwd, uint32_t* tw_inv, uint32_t* post_inv, uint32_t* post_a,
                             uint32_t* post_ci, uint32_t* post_h, uint32_t* pre_coset) {
    const size_t n = (size_t)1 << log_n;
    size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (k >= n) return;
    Fr w = Fr::from_const(FR_ROOT_2_32);
    for (unsigned i = log_n; i < FR_TWO_ADICITY; i++) w = w.sqr();  // w = primitive n-th root
    Fr g = Fr::from_const(FR_GENERATOR);
    Fr wk = fr_pow_u64(w, k);
    wk.store(tw_fwd + k * 8);
    Fr wi = (k == 0) ? Fr::one() : fr_pow_u64(w, n - k);
    wi.store(tw_inv + k * 8);
    // n^-1
    Fr nn = Fr::zero();
    nn.l[0] = (uint32_t)n;
    nn.l[1] = (uint32_t)(n >> 32);
    Fr ninv = nn.to_mont().inv();
    if (k == 0) ninv.store(post_inv);
    Fr gk = fr_pow_u64(g, k);
    gk.store(pre_coset + k * 8);
    (gk * ninv).store(post_a + k * 8);
    Fr gik = gk.inv();
    Fr ci = gik * ninv;
    ci.store(post_ci + k * 8);
    Fr zg = fr_pow_u64(g, n) - Fr::one();  // vanishing polynomial on the coset
    (ci * zg.inv()).from_mont().store(post_h + k * 8);
}

int ntt_domain_create(NttDomain& d, unsigned log_n, cudaStream_t st) {
    if (log_n > NTT_MAX_LOG) { set_error_detail("ntt: log_n = %u exceeds %u", log_n, NTT_MAX_LOG); return MP_ERR_UNSUPPORTED; }
    d.log_n = log_n;
    d.n = (size_t)1 << log_n;
    size_t bytes = d.n * 32;
    MP_TRY(d.tw_fwd.alloc(bytes));
    MP_TRY(d.tw_inv.alloc(bytes));
    MP_TRY(d.post_inv.alloc(32));
    MP_TRY(d.post_coset_a.alloc(bytes));
    MP_TRY(d.post_coset_inv.alloc(bytes));
    MP_TRY(d.post_h.alloc(bytes));
    MP_TRY(d.pre_coset.alloc(bytes));
    k_ntt_tables<<<div_up(d.n, 64), 64, 0, st>>>(log_n, d.tw_fwd.as<uint32_t>(), d.tw_inv.as<uint32_t>(), d.post_inv.as<uint32_t>(),
                                               d.post_coset_a.as<uint32_t>(), d.post_coset_inv.as<uint32_t>(),
                                               d.post_h.as<uint32_t>(), d.pre_coset.as<uint32_t>());
    MP_KERNEL_CHECK();
    return MP_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Montgomery conversion, R1CS evaluation, pointwise quotient numerator
// ---------------------------------------------------------------------------------------------------------
__global__ void k_fr_convert(const uint32_t* in, uint32_t* out, size_t n, int to_mont) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr x = Fr::load(in + i * 8);
    x = to_mont ? x.to_mont() : x.from_mont();
    x.store(out + i * 8);
}
int fr_to_mont(const void* in, void* out, size_t n, cudaStream_t st) {
    if (!n) return MP_OK;
    k_fr_convert<<<div_up(n, 256), 256, 0, st>>>((const uint32_t*)in, (uint32_t*)out, n, 1);
    MP_KERNEL_CHECK();
    return MP_OK;
}
int fr_from_mont(const void* in, void* out, size_t n, cudaStream_t st) {
    if (!n) return MP_OK;
    k_fr_convert<<<div_up(n, 256), 256, 0, st>>>((const uint32_t*)in, (uint32_t*)out, n, 0);
    MP_KERNEL_CHECK();
    return MP_OK;
}

struct SpmvArgs {
    const uint32_t* row_ptr[3];
    const uint32_t* col[3];
    const uint32_t* coeff[3];
};

// blockIdx.y = vector, blockIdx.z = matrix; thread = row of the padded domain
__global__ void k_r1cs_eval(SpmvArgs a, uint32_t K, uint32_t p, uint32_t m, const uint32_t* __restrict__ z, size_t z_stride, uint32_t* abc) {
    uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= m) return;
    const uint32_t v = blockIdx.y, mat = blockIdx.z;
    const uint32_t* zz = z + (size_t)v * z_stride * 8;
    Fr acc = Fr::zero();
    if (row < K) {
        uint32_t e0 = a.row_ptr[mat][row], e1 = a.row_ptr[mat][row + 1];
        for (uint32_t e = e0; e < e1; e++) {
            Fr c = Fr::load(a.coeff[mat] + (size_t)e * 8);
            Fr x = Fr::load(zz + (size_t)a.col[mat][e] * 8);
            acc = acc + c * x;
        }
    } else if (mat == 0 && row < K + p) {
        acc = Fr::load(zz + (size_t)(row - K) * 8);
    }
    acc.store(abc + (((size_t)v * 3 + mat) * m + row) * 8);
}

__global__ void k_r1cs_coeff_to_mont(uint32_t* coeff, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) Fr::load(coeff + i * 8).to_mont().store(coeff + i * 8);
}

int r1cs_upload(R1csDev& r, const mp_r1cs_view* v, cudaStream_t st) {
    r.p = v->num_instance;
    r.w = v->num_witness;
    r.K = v->num_constraints;
    const uint64_t* rp[3] = {v->a_row_ptr, v->b_row_ptr, v->c_row_ptr};
    const uint32_t* cl[3] = {v->a_col, v->b_col, v->c_col};
    const uint64_t* cf[3] = {v->a_coeff, v->b_coeff, v->c_coeff};
    for (int m = 0; m < 3; m++) {
        if (!rp[m]) return MP_ERR_INVALID_ARG;
        size_t nnz = rp[m][r.K];
        if (nnz >= (1ull << 32)) return MP_ERR_UNSUPPORTED;
        if (nnz && (!cl[m] || !cf[m])) return MP_ERR_INVALID_ARG;
        std::vector<uint32_t> rp32(r.K + 1);  // narrow row_ptr to u32
        for (size_t i = 0; i <= r.K; i++) {
            if (rp[m][i] > nnz || (i && rp[m][i] < rp[m][i - 1])) return MP_ERR_FORMAT;
            rp32[i] = (uint32_t)rp[m][i];
        }
        for (size_t e = 0; e < nnz; e++)
            if (cl[m][e] >= r.p + r.w) return MP_ERR_FORMAT;
        MP_TRY(r.row_ptr[m].alloc((r.K + 1) * 4));
        MP_TRY(r.col[m].alloc(nnz * 4));
        MP_TRY(r.coeff[m].alloc(nnz * 32));
        MP_CUDA_TRY(cudaMemcpyAsync(r.row_ptr[m].p, rp32.data(), (r.K + 1) * 4, cudaMemcpyHostToDevice, st));
        MP_CUDA_TRY(cudaStreamSynchronize(st));
        if (nnz) {
            MP_CUDA_TRY(cudaMemcpyAsync(r.col[m].p, cl[m], nnz * 4, cudaMemcpyHostToDevice, st));
            MP_CUDA_TRY(cudaMemcpyAsync(r.coeff[m].p, cf[m], nnz * 32, cudaMemcpyHostToDevice, st));
            k_r1cs_coeff_to_mont<<<div_up(nnz, 256), 256, 0, st>>>(r.coeff[m].as<uint32_t>(), nnz);
            MP_KERNEL_CHECK();
        }
    }
    MP_CUDA_TRY(cudaStreamSynchronize(st));
    return MP_OK;
}

int r1cs_eval(const R1csDev& r, const void* z_mont, size_t z_stride, size_t count, size_t m, void* abc, cudaStream_t st) {
    if (!count) return MP_OK;
    SpmvArgs a{};
    for (int i = 0; i < 3; i++) {
        a.row_ptr[i] = r.row_ptr[i].as<uint32_t>();
        a.col[i] = r.col[i].as<uint32_t>();
        a.coeff[i] = r.coeff[i].as<uint32_t>();
    }
    k_r1cs_eval<<<dim3(div_up(m, 128), (unsigned)count, 3), 128, 0, st>>>(a, (uint32_t)r.K, (uint32_t)r.p, (uint32_t)m,
                                                                       (const uint32_t*)z_mont, z_stride, (uint32_t*)abc);
    MP_KERNEL_CHECK();
    return MP_OK;
}

// s[v][i] = a[v][i] * b[v][i] - c[v][i]
__global__ void k_quotient_numerator(const uint32_t* __restrict__ abc, uint32_t* out, size_t m, size_t count) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t v = blockIdx.y;
    if (i >= m) return;
    const uint32_t* base = abc + v * 3 * m * 8;
    Fr a = Fr::load(base + i * 8), b = Fr::load(base + (m + i) * 8), c = Fr::load(base + (2 * m + i) * 8);
    (a * b - c).store(out + (v * m + i) * 8);
}

int witness_map_run(const NttDomain& d, void* abc, void* s1, void* s2, size_t count, void* out_h, size_t h_stride, cudaStream_t st) {
    if (!count) return MP_OK;
    const size_t m = d.n;
    // a, b, c: ifft then coset shift (g^i / m folded into the store) ...
    MP_TRY(ntt_run(d, true, abc, s1, s2, count * 3, nullptr, d.post_coset_a.p, nullptr, st));
    // ... then the forward transform gives the evaluations on the coset
    MP_TRY(ntt_run(d, false, s1, abc, s2, count * 3, nullptr, nullptr, nullptr, st));
    k_quotient_numerator<<<dim3(div_up(m, 256), (unsigned)count), 256, 0, st>>>((const uint32_t*)abc, (uint32_t*)s1, m, count);
    MP_KERNEL_CHECK();
    // coset_ifft with 1/(m Z(g)) g^-i and the Montgomery -> canonical conversion folded into the store
    MP_TRY(ntt_run_strided(d, true, s1, out_h, s2, count, m, h_stride, nullptr, d.post_h.p, nullptr, st, nullptr));
    return MP_OK;
}

}  // namespace mp

using namespace mp;

extern "C" int mp_ntt(int device, uint64_t* data, unsigned log_n, int inverse, int coset, float* out_ms) {
    if (!data) return MP_ERR_INVALID_ARG;
    MP_TRY(use_device(device));
    NttDomain d;
    MP_TRY(ntt_domain_create(d, log_n, 0));
    size_t bytes = d.n * 32;
    DevBuf a, b, c, c2;
    MP_TRY(a.alloc(bytes));
    MP_TRY(b.alloc(bytes));
    MP_TRY(c.alloc(bytes));
    if (log_n > 2 * NTT_MAX_SUB) MP_TRY(c2.alloc(bytes));
    MP_CUDA_TRY(cudaMemcpy(a.p, data, bytes, cudaMemcpyHostToDevice));
    MP_TRY(fr_to_mont(a.p, a.p, d.n, 0));
    EventTimer timer;
    MP_TRY(timer.start(0));
    const void* pre = (!inverse && coset) ? d.pre_coset.p : nullptr;
    const void* post = (inverse && coset) ? d.post_coset_inv.p : nullptr;
    const void* post_c = (inverse && !coset) ? d.post_inv.p : nullptr;
    MP_TRY(ntt_run_strided(d, inverse != 0, a.p, b.p, c.p, 1, d.n, d.n, pre, post, post_c, 0, c2.p));
    MP_TRY(timer.stop(0));
    MP_TRY(fr_from_mont(b.p, b.p, d.n, 0));
    MP_CUDA_TRY(cudaMemcpy(data, b.p, bytes, cudaMemcpyDeviceToHost));
    float ms = 0;
    MP_TRY(timer.elapsed_ms(&ms));
    if (out_ms) *out_ms = ms;
    return MP_OK;
}
