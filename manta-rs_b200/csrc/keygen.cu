// Key generation on the device: QAP evaluation at the toxic waste, the fixed-base multiplications of
// `circuit_specific_setup`, and the phase-2 `initialize` of the trusted setup with its four group-valued inverse FFTs.
//
// Replaces, for BLS12-381:
//   * ark-groth16 0.3 `generate_parameters` as reached from manta-crypto/src/arkworks/groth16.rs:570-586 (`compile`), with the
//     standard generators and a caller-supplied trapdoor (tau, alpha, beta, gamma, delta): `mp_keygen`;
//   * manta-trusted-setup/src/groth16/mpc.rs:355-431 (`initialize`): h_query[i] = tau^(i+m) G - tau^i G, ifft over G1 / G2 of
//     the tau, alpha tau, beta tau powers (ark-poly `domain.ifft` on group elements), the sparse accumulation
//     `specialize_to_phase_2` (:245-293) with `add_dummy_constraints` (:295-312), batch normalisation: `mp_mpc_initialize`;
//   * the group-valued radix-2 transform on its own: `mp_group_ntt`.
// Outputs are `ProvingContext` bytes (groth16.rs:290-303).  Only affine values are observable, so the butterfly order,
// the XYZZ coordinates and the per-thread double-and-add scalar multiplications give the reference's bytes.
#include <algorithm>
#include <vector>

#include "msm.cuh"

namespace mp {

// ---------------------------------------------------------------------------------------------------------
// column-major (CSC) copy of the constraint matrices: the QAP polynomials and the phase-2 accumulation are sums per VARIABLE
// ---------------------------------------------------------------------------------------------------------
struct CscDev {
    uint64_t n = 0, p = 0, K = 0;
    DevBuf col_ptr[3], row[3], coeff[3];   // coeff canonical (4 x u64 each)
};
struct CscArgs {
    const uint32_t* col_ptr[3];
    const uint32_t* row[3];
    const uint32_t* coeff[3];
};
static CscArgs csc_args(const CscDev& c) {
    CscArgs a{};
    for (int m = 0; m < 3; m++) {
        a.col_ptr[m] = c.col_ptr[m].as<uint32_t>();
        a.row[m] = c.row[m].as<uint32_t>();
        a.coeff[m] = c.coeff[m].as<uint32_t>();
    }
    return a;
}

static int csc_upload(CscDev& c, const mp_r1cs_view* v, cudaStream_t st) {
    c.p = v->num_instance;
    c.K = v->num_constraints;
    c.n = v->num_instance + v->num_witness;
    if (c.p == 0 || c.n >= (1u << 26) || c.K >= (1u << 28)) { set_error_detail("bad R1CS shape"); return MP_ERR_INVALID_ARG; }
    const uint64_t* rp[3] = {v->a_row_ptr, v->b_row_ptr, v->c_row_ptr};
    const uint32_t* cl[3] = {v->a_col, v->b_col, v->c_col};
    const uint64_t* cf[3] = {v->a_coeff, v->b_coeff, v->c_coeff};
    for (int m = 0; m < 3; m++) {
        if (!rp[m]) return MP_ERR_INVALID_ARG;
        const size_t nnz = rp[m][c.K];
        if (nnz >= (1ull << 31)) return MP_ERR_UNSUPPORTED;
        if (nnz && (!cl[m] || !cf[m])) return MP_ERR_INVALID_ARG;
        std::vector<uint32_t> ptr(c.n + 1, 0), rows(std::max<size_t>(nnz, 1));
        std::vector<uint64_t> coeffs(std::max<size_t>(nnz, 1) * 4);
        for (size_t i = 0; i < c.K; i++)
            if (rp[m][i + 1] < rp[m][i] || rp[m][i + 1] > nnz) return MP_ERR_FORMAT;
        for (size_t e = 0; e < nnz; e++) {
            if (cl[m][e] >= c.n) return MP_ERR_FORMAT;
            ptr[cl[m][e] + 1]++;
        }
        for (size_t i = 0; i < c.n; i++) ptr[i + 1] += ptr[i];
        std::vector<uint32_t> cur(ptr.begin(), ptr.end() - 1);
        for (size_t r = 0; r < c.K; r++)
            for (uint64_t e = rp[m][r]; e < rp[m][r + 1]; e++) {
                const uint32_t pos = cur[cl[m][e]]++;
                rows[pos] = (uint32_t)r;
                memcpy(&coeffs[4 * (size_t)pos], cf[m] + 4 * e, 32);
            }
        MP_TRY(c.col_ptr[m].alloc((c.n + 1) * 4));
        MP_TRY(c.row[m].alloc(rows.size() * 4));
        MP_TRY(c.coeff[m].alloc(coeffs.size() * 8));
        MP_CUDA_TRY(cudaMemcpyAsync(c.col_ptr[m].p, ptr.data(), (c.n + 1) * 4, cudaMemcpyHostToDevice, st));
        MP_CUDA_TRY(cudaMemcpyAsync(c.row[m].p, rows.data(), rows.size() * 4, cudaMemcpyHostToDevice, st));
        MP_CUDA_TRY(cudaMemcpyAsync(c.coeff[m].p, coeffs.data(), coeffs.size() * 8, cudaMemcpyHostToDevice, st));
        MP_CUDA_TRY(cudaStreamSynchronize(st));   // the staging vectors die with this iteration
    }
    return MP_OK;
}

MP_DEV Fr fr_root_of_unity(unsigned log_m) {
    Fr w = Fr::from_const(FR_ROOT_2_32);
    for (unsigned i = log_m; i < FR_TWO_ADICITY; i++) w = w.sqr();
    return w;
}
MP_DEV Fr fr_from_u64(uint64_t v) {
    Fr x = Fr::zero();
    x.l[0] = (uint32_t)v;
    x.l[1] = (uint32_t)(v >> 32);
    return x.to_mont();
}

// ---------------------------------------------------------------------------------------------------------
// QAP at tau
// ---------------------------------------------------------------------------------------------------------
// consts (Montgomery): 0 tau, 1 alpha, 2 beta, 3 gamma, 4 delta, 5 1/gamma, 6 1/delta, 7 Z(tau) = tau^m - 1
__global__ void k_trapdoor_consts(const uint32_t* __restrict__ trap_canon, unsigned log_m, uint32_t* consts, uint32_t* bad) {
    if (threadIdx.x || blockIdx.x) return;
    Fr t[5];
    for (int i = 0; i < 5; i++) {
        t[i] = Fr::load(trap_canon + 8 * i);
        uint32_t tmp[Fr::N];
        if (!Fr::sub_raw(tmp, t[i].l, FrParams::mod()) || t[i].is_zero()) atomicOr(bad, 1u);   // every trapdoor element canonical and non-zero
        t[i] = t[i].to_mont();
        t[i].store(consts + 8 * i);
    }
    t[3].inv().store(consts + 8 * 5);
    t[4].inv().store(consts + 8 * 6);
    Fr zt = t[0];
    for (unsigned i = 0; i < log_m; i++) zt = zt.sqr();
    zt = zt - Fr::one();
    if (zt.is_zero()) atomicOr(bad, 2u);   // tau inside the domain: not a valid setup
    zt.store(consts + 8 * 7);
}

// L_j(tau) = Z(tau) / m * w^j / (tau - w^j)
__global__ void __launch_bounds__(64) k_lagrange_at_tau(unsigned log_m, const uint32_t* __restrict__ consts, uint32_t* L) {
    const size_t m = (size_t)1 << log_m, j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (j >= m) return;
    const Fr tau = Fr::load(consts), zt = Fr::load(consts + 8 * 7);
    const Fr wj = fr_root_of_unity(log_m).pow_u64(j);
    const Fr minv = fr_from_u64(m).inv();
    (zt * minv * wj * (tau - wj).inv()).store(L + j * 8);
}

// uvw[mat][i] = sum over the entries of column i of matrix mat: coeff * L[row]   (+ L[K + i] for the A polynomial of a public variable)
__global__ void __launch_bounds__(128) k_qap_columns(CscArgs a, uint32_t n, uint32_t p, uint32_t K, const uint32_t* __restrict__ L, uint32_t* uvw) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, mat = blockIdx.y;
    if (i >= n) return;
    Fr acc = Fr::zero();
    for (uint32_t e = a.col_ptr[mat][i]; e < a.col_ptr[mat][i + 1]; e++)
        acc = acc + Fr::load(a.coeff[mat] + (size_t)e * 8).to_mont() * Fr::load(L + (size_t)a.row[mat][e] * 8);
    if (mat == 0 && i < p) acc = acc + Fr::load(L + (size_t)(K + i) * 8);
    acc.store(uvw + ((size_t)mat * n + i) * 8);
}

// Discrete logs of every key element, canonical, in file order:
//   G1: alpha | (beta u_i + alpha v_i + w_i) / gamma (i < p) | beta | delta | u (n) | v (n) | tau^k Z(tau) / delta (k < h_len) | (...) / delta (i >= p)
//   G2: beta | gamma | delta | v (n)
__global__ void __launch_bounds__(128) k_key_scalars(uint32_t n, uint32_t p, uint32_t h_len, const uint32_t* __restrict__ consts,
                                                    const uint32_t* __restrict__ uvw, uint32_t* g1, uint32_t* g2) {
    const size_t n1 = (size_t)3 + p + 2 * (size_t)n + h_len + (n - p), n2 = (size_t)3 + n;
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= n1 + n2) return;
    const Fr alpha = Fr::load(consts + 8), beta = Fr::load(consts + 16);
    auto abc = [&](size_t i) { return beta * Fr::load(uvw + i * 8) + alpha * Fr::load(uvw + ((size_t)n + i) * 8) + Fr::load(uvw + (2 * (size_t)n + i) * 8); };
    Fr v;
    uint32_t* out;
    if (t < n1) {
        out = g1 + t * 8;
        size_t k = t;
        if (k == 0) v = alpha;
        else if (k < 1 + (size_t)p) v = abc(k - 1) * Fr::load(consts + 8 * 5);
        else if (k == 1 + (size_t)p) v = beta;
        else if (k == 2 + (size_t)p) v = Fr::load(consts + 8 * 4);
        else if ((k -= 3 + (size_t)p) < n) v = Fr::load(uvw + k * 8);
        else if ((k -= n) < n) v = Fr::load(uvw + ((size_t)n + k) * 8);
        else if ((k -= n) < h_len) v = Fr::load(consts + 8 * 7) * Fr::load(consts + 8 * 6) * Fr::load(consts).pow_u64(k);
        else v = abc(p + (k - h_len)) * Fr::load(consts + 8 * 6);
    } else {
        const size_t k = t - n1;
        out = g2 + k * 8;
        v = k == 0 ? beta : (k == 1 ? Fr::load(consts + 8 * 3) : (k == 2 ? Fr::load(consts + 8 * 4) : Fr::load(uvw + ((size_t)n + (k - 3)) * 8)));
    }
    v.from_mont().store(out);
}

// ---------------------------------------------------------------------------------------------------------
// group-valued radix-2 transform (ark-poly `Radix2EvaluationDomain::{fft, ifft}` over `DomainCoeff` = curve points)
// ---------------------------------------------------------------------------------------------------------
template <class F>
MP_COLD XYZZ<F> point_mul(const XYZZ<F>& p, const Fr& k_canon) {
    XYZZ<F> r = XYZZ<F>::inf();
    int bit = 254;
    while (bit >= 0 && !((k_canon.l[bit >> 5] >> (bit & 31)) & 1)) bit--;
    for (; bit >= 0; bit--) {
        r = r.dbl();
        if ((k_canon.l[bit >> 5] >> (bit & 31)) & 1) r = r.add(p);
    }
    return r;
}

// out[bitrev(i)] = in[i] as XYZZ
template <class F>
__global__ void __launch_bounds__(128) k_group_bitrev(const uint32_t* __restrict__ in, XYZZ<F>* out, unsigned log_n) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= ((size_t)1 << log_n)) return;
    const size_t r = log_n ? (__brevll((unsigned long long)i) >> (64 - log_n)) : 0;
    XYZZ<F>::from_affine(Affine<F>::load(in + i * Affine<F>::WORDS)).store(out + r);
}

// One decimation-in-time level: blocks of len = 2^(s+1); (u, v) -> (u + w^e v, u - w^e v), e = j * n / len
template <class F>
__global__ void __launch_bounds__(64) k_group_butterfly(XYZZ<F>* a, unsigned log_n, unsigned s, int inverse) {
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x, half = (size_t)1 << s;
    if (t >= ((size_t)1 << log_n) / 2) return;
    const size_t j = t & (half - 1), k = (t >> s) << (s + 1);
    XYZZ<F> u = XYZZ<F>::load(a + k + j), v = XYZZ<F>::load(a + k + j + half);
    if (j) {
        uint64_t e = (uint64_t)j << (log_n - s - 1);
        if (inverse) e = ((uint64_t)1 << log_n) - e;
        v = point_mul(v, fr_root_of_unity(log_n).pow_u64(e).from_mont());
    }
    u.add(v).store(a + k + j);
    u.add(v.neg()).store(a + k + j + half);
}

// affine output, times 1 / n for the inverse transform
template <class F>
__global__ void __launch_bounds__(64) k_group_finish(const XYZZ<F>* __restrict__ a, uint32_t* out, unsigned log_n, int inverse) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= ((size_t)1 << log_n)) return;
    XYZZ<F> p = XYZZ<F>::load(a + i);
    if (inverse && log_n) p = point_mul(p, fr_from_u64((uint64_t)1 << log_n).inv().from_mont());
    p.to_affine().store(out + i * Affine<F>::WORDS);
}

// d_points: n = 2^log_n affine Montgomery points, transformed in place; d_work: n XYZZ<F>
template <class F>
static int group_ntt_dev(void* d_points, void* d_work, unsigned log_n, bool inverse, cudaStream_t st) {
    const size_t n = (size_t)1 << log_n;
    k_group_bitrev<F><<<div_up(n, 128), 128, 0, st>>>((const uint32_t*)d_points, (XYZZ<F>*)d_work, log_n);
    MP_KERNEL_CHECK();
    for (unsigned s = 0; s < log_n; s++) {
        k_group_butterfly<F><<<div_up(n / 2, 64), 64, 0, st>>>((XYZZ<F>*)d_work, log_n, s, inverse ? 1 : 0);
        MP_KERNEL_CHECK();
    }
    k_group_finish<F><<<div_up(n, 64), 64, 0, st>>>((const XYZZ<F>*)d_work, (uint32_t*)d_points, log_n, inverse ? 1 : 0);
    MP_KERNEL_CHECK();
    return MP_OK;
}

// ---------------------------------------------------------------------------------------------------------
// phase-2 initialisation (mpc.rs:245-312, 355-431)
// ---------------------------------------------------------------------------------------------------------
// h_query[i] = tau^(i + m) G - tau^i G
__global__ void __launch_bounds__(64) k_mpc_h_query(const uint32_t* __restrict__ tau_g1, uint32_t m, uint32_t* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    constexpr int AW = Affine<Fq>::WORDS;
    XYZZ<Fq> hi = XYZZ<Fq>::from_affine(Affine<Fq>::load(tau_g1 + (size_t)(i + m) * AW));
    hi.add_mixed_cold(Affine<Fq>::load(tau_g1 + (size_t)i * AW).neg()).to_affine().store(out + (size_t)i * AW);
}

template <class F>
MP_DEV XYZZ<F> mpc_column_sum(const CscArgs& a, int mat, uint32_t i, const uint32_t* lag) {
    constexpr int AW = Affine<F>::WORDS;
    XYZZ<F> acc = XYZZ<F>::inf();
    for (uint32_t e = a.col_ptr[mat][i]; e < a.col_ptr[mat][i + 1]; e++) {
        const XYZZ<F> pt = XYZZ<F>::from_affine(Affine<F>::load(lag + (size_t)a.row[mat][e] * AW));
        acc = acc.add(point_mul(pt, Fr::load(a.coeff[mat] + (size_t)e * 8)));
    }
    return acc;
}

// kind 0: a_g1, 1: b_g1, 2: ext (the cross terms, public ones first); thread = (variable, kind)
__global__ void __launch_bounds__(64) k_mpc_specialize_g1(CscArgs a, uint32_t n, uint32_t p, uint32_t K, const uint32_t* __restrict__ tau_l,
                                                         const uint32_t* __restrict__ alpha_l, const uint32_t* __restrict__ beta_l, uint32_t* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, kind = blockIdx.y;
    if (i >= n) return;
    constexpr int AW = Affine<Fq>::WORDS;
    XYZZ<Fq> acc;
    if (kind == 0) {
        acc = mpc_column_sum<Fq>(a, 0, i, tau_l);
        if (i < p) acc = acc.add_mixed_cold(Affine<Fq>::load(tau_l + (size_t)(K + i) * AW));     // add_dummy_constraints
    } else if (kind == 1) {
        acc = mpc_column_sum<Fq>(a, 1, i, tau_l);
    } else {
        acc = mpc_column_sum<Fq>(a, 0, i, beta_l).add(mpc_column_sum<Fq>(a, 1, i, alpha_l)).add(mpc_column_sum<Fq>(a, 2, i, tau_l));
        if (i < p) acc = acc.add_mixed_cold(Affine<Fq>::load(beta_l + (size_t)(K + i) * AW));
    }
    acc.to_affine().store(out + ((size_t)kind * n + i) * AW);
}
__global__ void __launch_bounds__(64) k_mpc_specialize_g2(CscArgs a, uint32_t n, const uint32_t* __restrict__ tau_l2, uint32_t* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    mpc_column_sum<Fq2>(a, 1, i, tau_l2).to_affine().store(out + (size_t)i * Affine<Fq2>::WORDS);
}

static __global__ void k_generators(uint32_t* g1, uint32_t* g2) {
    if (threadIdx.x || blockIdx.x) return;
    for (int c = 0; c < 2; c++) for (int k = 0; k < 12; k++) g1[c * 12 + k] = G1_GEN[c][k];
    for (int c = 0; c < 4; c++) for (int k = 0; k < 12; k++) g2[c * 12 + k] = G2_GEN[c][k];
}

static size_t pk_bytes(uint64_t n, uint64_t p, uint64_t h_len) {
    return MP_G1_BYTES + 3 * MP_G2_BYTES + 8 + p * MP_G1_BYTES + 2 * MP_G1_BYTES + 8 + n * MP_G1_BYTES + 8 + n * MP_G1_BYTES + 8 +
           n * MP_G2_BYTES + 8 + h_len * MP_G1_BYTES + 8 + (n - p) * MP_G1_BYTES;
}

// Writes the `ProvingContext` encoding from ark-form point blocks on the HOST.
struct PkPieces {
    const uint8_t *alpha_g1, *beta_g2, *gamma_g2, *delta_g2, *gamma_abc, *beta_g1, *delta_g1, *a, *b1, *b2, *h, *l;
};
static void pk_write(uint8_t* out, const PkPieces& k, uint64_t n, uint64_t p, uint64_t h_len) {
    auto put = [&](const uint8_t* src, size_t bytes) { memcpy(out, src, bytes); out += bytes; };
    auto vec = [&](const uint8_t* src, uint64_t cnt, size_t elem) { memcpy(out, &cnt, 8); out += 8; put(src, cnt * elem); };
    put(k.alpha_g1, MP_G1_BYTES);
    put(k.beta_g2, MP_G2_BYTES);
    put(k.gamma_g2, MP_G2_BYTES);
    put(k.delta_g2, MP_G2_BYTES);
    vec(k.gamma_abc, p, MP_G1_BYTES);
    put(k.beta_g1, MP_G1_BYTES);
    put(k.delta_g1, MP_G1_BYTES);
    vec(k.a, n, MP_G1_BYTES);
    vec(k.b1, n, MP_G1_BYTES);
    vec(k.b2, n, MP_G2_BYTES);
    vec(k.h, h_len, MP_G1_BYTES);
    vec(k.l, n - p, MP_G1_BYTES);
}

static unsigned domain_log(uint64_t K, uint64_t p) {
    unsigned lg = 0;
    while (((uint64_t)1 << lg) < K + p) lg++;
    return lg;
}

}  // namespace mp

using namespace mp;

extern "C" {

int mp_keygen(int device, const mp_r1cs_view* r1cs, const uint64_t* trapdoor, uint64_t h_len, uint8_t* out_pk, size_t out_cap, size_t* out_len) {
    if (!r1cs || !trapdoor || !out_len) return MP_ERR_INVALID_ARG;
    const uint64_t n = r1cs->num_instance + r1cs->num_witness, p = r1cs->num_instance;
    const unsigned log_m = domain_log(r1cs->num_constraints, p);
    if (log_m > 28) return MP_ERR_UNSUPPORTED;
    const uint64_t m = (uint64_t)1 << log_m;
    if (h_len == 0) h_len = m - 1;   // ark's generator; m for MPC-style keys
    if (h_len > m) return MP_ERR_INVALID_ARG;
    *out_len = pk_bytes(n, p, h_len);
    if (!out_pk || out_cap < *out_len) return out_pk ? MP_ERR_INVALID_ARG : MP_OK;   // size query
    MP_TRY(use_device(device));
    cudaStream_t st = 0;
    CscDev csc;
    MP_TRY(csc_upload(csc, r1cs, st));
    const size_t n1 = 3 + p + 2 * n + h_len + (n - p), n2 = 3 + n;
    DevBuf d_trap, d_consts, d_bad, d_L, d_uvw, d_s1, d_s2, d_p1, d_p2;
    MP_TRY(d_trap.alloc(5 * 32));
    MP_TRY(d_consts.alloc(8 * 32));
    MP_TRY(d_bad.alloc(4));
    MP_TRY(d_L.alloc(m * 32));
    MP_TRY(d_uvw.alloc(3 * n * 32));
    MP_TRY(d_s1.alloc(n1 * 32));
    MP_TRY(d_s2.alloc(n2 * 32));
    MP_TRY(d_p1.alloc(n1 * MP_G1_BYTES));
    MP_TRY(d_p2.alloc(n2 * MP_G2_BYTES));
    MP_CUDA_TRY(cudaMemcpyAsync(d_trap.p, trapdoor, 5 * 32, cudaMemcpyHostToDevice, st));
    MP_CUDA_TRY(cudaMemsetAsync(d_bad.p, 0, 4, st));
    k_trapdoor_consts<<<1, 1, 0, st>>>(d_trap.as<uint32_t>(), log_m, d_consts.as<uint32_t>(), d_bad.as<uint32_t>());
    MP_KERNEL_CHECK();
    uint32_t bad = 0;
    MP_CUDA_TRY(cudaMemcpyAsync(&bad, d_bad.p, 4, cudaMemcpyDeviceToHost, st));
    MP_CUDA_TRY(cudaStreamSynchronize(st));
    if (bad) { set_error_detail("keygen: trapdoor element zero / not canonical, or tau inside the evaluation domain"); return MP_ERR_INVALID_ARG; }
    k_lagrange_at_tau<<<div_up(m, 64), 64, 0, st>>>(log_m, d_consts.as<uint32_t>(), d_L.as<uint32_t>());
    MP_KERNEL_CHECK();
    k_qap_columns<<<dim3(div_up(n, 128), 3), 128, 0, st>>>(csc_args(csc), (uint32_t)n, (uint32_t)p, (uint32_t)csc.K, d_L.as<uint32_t>(), d_uvw.as<uint32_t>());
    MP_KERNEL_CHECK();
    k_key_scalars<<<div_up(n1 + n2, 128), 128, 0, st>>>((uint32_t)n, (uint32_t)p, (uint32_t)h_len, d_consts.as<uint32_t>(), d_uvw.as<uint32_t>(),
                                                       d_s1.as<uint32_t>(), d_s2.as<uint32_t>());
    MP_KERNEL_CHECK();
    MP_TRY(msm_fixed_base_dev_g1(d_s1.p, n1, d_p1.p, st));
    MP_TRY(msm_fixed_base_dev_g2(d_s2.p, n2, d_p2.p, st));
    MP_TRY(points_to_ark_g1(d_p1.p, d_p1.p, n1, st));
    MP_TRY(points_to_ark_g2(d_p2.p, d_p2.p, n2, st));
    std::vector<uint8_t> g1(n1 * MP_G1_BYTES), g2(n2 * MP_G2_BYTES);
    MP_CUDA_TRY(cudaMemcpyAsync(g1.data(), d_p1.p, g1.size(), cudaMemcpyDeviceToHost, st));
    MP_CUDA_TRY(cudaMemcpyAsync(g2.data(), d_p2.p, g2.size(), cudaMemcpyDeviceToHost, st));
    MP_CUDA_TRY(cudaStreamSynchronize(st));
    PkPieces k{};
    const uint8_t* q = g1.data();
    auto take = [&](size_t cnt) { const uint8_t* r = q; q += cnt * MP_G1_BYTES; return r; };
    k.alpha_g1 = take(1); k.gamma_abc = take(p); k.beta_g1 = take(1); k.delta_g1 = take(1);
    k.a = take(n); k.b1 = take(n); k.h = take(h_len); k.l = take(n - p);
    k.beta_g2 = g2.data(); k.gamma_g2 = g2.data() + MP_G2_BYTES; k.delta_g2 = g2.data() + 2 * MP_G2_BYTES; k.b2 = g2.data() + 3 * MP_G2_BYTES;
    pk_write(out_pk, k, n, p, h_len);
    return MP_OK;
}

int mp_group_ntt(int device, int group, uint8_t* points, unsigned log_n, int inverse) {
    if (!points || (group != 1 && group != 2) || log_n > 24) return MP_ERR_INVALID_ARG;
    MP_TRY(use_device(device));
    const size_t n = (size_t)1 << log_n, pb = group == 1 ? MP_G1_BYTES : MP_G2_BYTES;
    DevBuf d_pts, d_work;
    MP_TRY(d_pts.alloc(n * pb));
    MP_TRY(d_work.alloc(n * pb * 2));
    MP_CUDA_TRY(cudaMemcpy(d_pts.p, points, n * pb, cudaMemcpyHostToDevice));
    if (group == 1) {
        MP_TRY(points_from_ark_g1(d_pts.p, d_pts.p, n, 0));
        MP_TRY(group_ntt_dev<Fq>(d_pts.p, d_work.p, log_n, inverse != 0, 0));
        MP_TRY(points_to_ark_g1(d_pts.p, d_pts.p, n, 0));
    } else {
        MP_TRY(points_from_ark_g2(d_pts.p, d_pts.p, n, 0));
        MP_TRY(group_ntt_dev<Fq2>(d_pts.p, d_work.p, log_n, inverse != 0, 0));
        MP_TRY(points_to_ark_g2(d_pts.p, d_pts.p, n, 0));
    }
    MP_CUDA_TRY(cudaMemcpy(points, d_pts.p, n * pb, cudaMemcpyDeviceToHost));
    return MP_OK;
}

int mp_mpc_initialize(int device, const mp_r1cs_view* r1cs, const uint8_t* tau_powers_g1, size_t n_tau_g1, const uint8_t* tau_powers_g2,
                      const uint8_t* alpha_tau_powers_g1, const uint8_t* beta_tau_powers_g1, const uint8_t* beta_g2, uint8_t* out_pk,
                      size_t out_cap, size_t* out_len) {
    if (!r1cs || !out_len) return MP_ERR_INVALID_ARG;
    const uint64_t n = r1cs->num_instance + r1cs->num_witness, p = r1cs->num_instance, K = r1cs->num_constraints;
    const unsigned log_m = domain_log(K, p);
    if (log_m > 24) return MP_ERR_UNSUPPORTED;
    const uint64_t m = (uint64_t)1 << log_m;
    *out_len = pk_bytes(n, p, m);
    if (!out_pk) return MP_OK;   // size query
    if (out_cap < *out_len || !tau_powers_g1 || !tau_powers_g2 || !alpha_tau_powers_g1 || !beta_tau_powers_g1 || !beta_g2) return MP_ERR_INVALID_ARG;
    if (n_tau_g1 < 2 * m) { set_error_detail("mpc initialize: %zu tau powers in G1, the domain of size %llu needs %llu", n_tau_g1, (unsigned long long)m, (unsigned long long)(2 * m)); return MP_ERR_INVALID_ARG; }
    MP_TRY(use_device(device));
    cudaStream_t st = 0;
    CscDev csc;
    MP_TRY(csc_upload(csc, r1cs, st));
    DevBuf d_tau1, d_tau2, d_alpha, d_beta, d_work, d_h, d_g1out, d_g2out, d_gen1, d_gen2;
    MP_TRY(d_tau1.alloc(2 * m * MP_G1_BYTES));
    MP_TRY(d_tau2.alloc(m * MP_G2_BYTES));
    MP_TRY(d_alpha.alloc(m * MP_G1_BYTES));
    MP_TRY(d_beta.alloc(m * MP_G1_BYTES));
    MP_TRY(d_work.alloc(m * 2 * MP_G2_BYTES));
    MP_TRY(d_h.alloc(m * MP_G1_BYTES));
    MP_TRY(d_g1out.alloc(3 * n * MP_G1_BYTES));
    MP_TRY(d_g2out.alloc(n * MP_G2_BYTES));
    MP_TRY(d_gen1.alloc(MP_G1_BYTES));
    MP_TRY(d_gen2.alloc(MP_G2_BYTES));
    MP_CUDA_TRY(cudaMemcpyAsync(d_tau1.p, tau_powers_g1, 2 * m * MP_G1_BYTES, cudaMemcpyHostToDevice, st));
    MP_CUDA_TRY(cudaMemcpyAsync(d_tau2.p, tau_powers_g2, m * MP_G2_BYTES, cudaMemcpyHostToDevice, st));
    MP_CUDA_TRY(cudaMemcpyAsync(d_alpha.p, alpha_tau_powers_g1, m * MP_G1_BYTES, cudaMemcpyHostToDevice, st));
    MP_CUDA_TRY(cudaMemcpyAsync(d_beta.p, beta_tau_powers_g1, m * MP_G1_BYTES, cudaMemcpyHostToDevice, st));
    MP_TRY(points_from_ark_g1(d_tau1.p, d_tau1.p, 2 * m, st));
    MP_TRY(points_from_ark_g2(d_tau2.p, d_tau2.p, m, st));
    MP_TRY(points_from_ark_g1(d_alpha.p, d_alpha.p, m, st));
    MP_TRY(points_from_ark_g1(d_beta.p, d_beta.p, m, st));
    // alpha_1 and beta_1 are the zeroth powers: keep their ark bytes before the transforms overwrite the arrays
    k_mpc_h_query<<<div_up(m, 64), 64, 0, st>>>(d_tau1.as<uint32_t>(), (uint32_t)m, d_h.as<uint32_t>());
    MP_KERNEL_CHECK();
    MP_TRY(group_ntt_dev<Fq>(d_tau1.p, d_work.p, log_m, true, st));     // the first m powers only (ark resizes to the domain)
    MP_TRY(group_ntt_dev<Fq2>(d_tau2.p, d_work.p, log_m, true, st));
    MP_TRY(group_ntt_dev<Fq>(d_alpha.p, d_work.p, log_m, true, st));
    MP_TRY(group_ntt_dev<Fq>(d_beta.p, d_work.p, log_m, true, st));
    k_mpc_specialize_g1<<<dim3(div_up(n, 64), 3), 64, 0, st>>>(csc_args(csc), (uint32_t)n, (uint32_t)p, (uint32_t)K, d_tau1.as<uint32_t>(),
                                                              d_alpha.as<uint32_t>(), d_beta.as<uint32_t>(), d_g1out.as<uint32_t>());
    MP_KERNEL_CHECK();
    k_mpc_specialize_g2<<<div_up(n, 64), 64, 0, st>>>(csc_args(csc), (uint32_t)n, d_tau2.as<uint32_t>(), d_g2out.as<uint32_t>());
    MP_KERNEL_CHECK();
    k_generators<<<1, 1, 0, st>>>(d_gen1.as<uint32_t>(), d_gen2.as<uint32_t>());
    MP_KERNEL_CHECK();
    MP_TRY(points_to_ark_g1(d_h.p, d_h.p, m, st));
    MP_TRY(points_to_ark_g1(d_g1out.p, d_g1out.p, 3 * n, st));
    MP_TRY(points_to_ark_g2(d_g2out.p, d_g2out.p, n, st));
    MP_TRY(points_to_ark_g1(d_gen1.p, d_gen1.p, 1, st));
    MP_TRY(points_to_ark_g2(d_gen2.p, d_gen2.p, 1, st));
    std::vector<uint8_t> h(m * MP_G1_BYTES), g1(3 * n * MP_G1_BYTES), g2(n * MP_G2_BYTES), gen1(MP_G1_BYTES), gen2(MP_G2_BYTES);
    MP_CUDA_TRY(cudaMemcpyAsync(h.data(), d_h.p, h.size(), cudaMemcpyDeviceToHost, st));
    MP_CUDA_TRY(cudaMemcpyAsync(g1.data(), d_g1out.p, g1.size(), cudaMemcpyDeviceToHost, st));
    MP_CUDA_TRY(cudaMemcpyAsync(g2.data(), d_g2out.p, g2.size(), cudaMemcpyDeviceToHost, st));
    MP_CUDA_TRY(cudaMemcpyAsync(gen1.data(), d_gen1.p, gen1.size(), cudaMemcpyDeviceToHost, st));
    MP_CUDA_TRY(cudaMemcpyAsync(gen2.data(), d_gen2.p, gen2.size(), cudaMemcpyDeviceToHost, st));
    MP_CUDA_TRY(cudaStreamSynchronize(st));
    PkPieces k{};
    k.alpha_g1 = alpha_tau_powers_g1;          // alpha tau^0 G
    k.beta_g2 = beta_g2;
    k.gamma_g2 = gen2.data();                  // mpc.rs:417-424: gamma = delta = 1
    k.delta_g2 = gen2.data();
    k.gamma_abc = g1.data() + 2 * n * MP_G1_BYTES;                    // ext[0 .. p)
    k.beta_g1 = beta_tau_powers_g1;            // beta tau^0 G
    k.delta_g1 = gen1.data();
    k.a = g1.data();
    k.b1 = g1.data() + n * MP_G1_BYTES;
    k.b2 = g2.data();
    k.h = h.data();
    k.l = g1.data() + (2 * n + p) * MP_G1_BYTES;                      // ext[p .. n)
    pk_write(out_pk, k, n, p, m);
    return MP_OK;
}

}  // extern "C"
