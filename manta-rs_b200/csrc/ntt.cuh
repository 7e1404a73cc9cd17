// Internal interface of the Fr number-theoretic transform and the R1CS -> QAP witness map (ntt.cu).
//
// Replaces ark-poly 0.3 `Radix2EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place` and
// ark-groth16 0.3 `R1CStoQAP::witness_map`.
#pragma once
#include "common.cuh"

namespace mp {

struct NttDomain {
    unsigned log_n = 0;
    size_t n = 0;
    DevBuf tw_fwd;   // omega^k, k < n       (Montgomery)
    DevBuf tw_inv;   // omega^-k
    DevBuf post_inv;       // 1/n                        (n copies not needed: single element, Montgomery)
    DevBuf post_coset_a;   // g^i / n                    (ifft immediately followed by coset shift)
    DevBuf post_coset_inv; // g^-i / n                   (coset_ifft), Montgomery
    DevBuf post_h;         // g^-i / (n * (g^n - 1)) in CANONICAL form: output of the last transform is canonical
    DevBuf pre_coset;      // g^i (coset_fft of a natural-order input)
};
int ntt_domain_create(NttDomain& d, unsigned log_n, cudaStream_t st);

// out[v][k] = sum_i (pre ? pre[i] : 1) * in[v][i] * w^(ik), times post[k] (table) or post_const[0] when given.
// `count` vectors of n elements, contiguous; in, out, tmp must be distinct buffers of count*n elements.
int ntt_run(const NttDomain& d, bool inverse, const void* in, void* out, void* tmp, size_t count,
            const void* pre_table, const void* post_table, const void* post_const, cudaStream_t st);

// General form: vectors `in_stride` / `out_stride` elements apart.  Transforms above 2^20 points take a third pass and need
// `tmp2`, a second scratch buffer of count * n elements (nullptr otherwise).
int ntt_run_strided(const NttDomain& d, bool inverse, const void* in, void* out, void* tmp, size_t count, size_t in_stride,
                    size_t out_stride, const void* pre_table, const void* post_table, const void* post_const, cudaStream_t st, void* tmp2);

// CSR matrices on the device, coefficients in Montgomery form.
struct R1csDev {
    uint64_t p = 0, w = 0, K = 0;
    DevBuf row_ptr[3], col[3], coeff[3];
};
int r1cs_upload(R1csDev& r, const mp_r1cs_view* v, cudaStream_t st);

// abc[v][3][m] (Montgomery): rows of A z, B z, C z, plus a[K + j] = z_j for j < p, zero padded to m.
// z_mont: [count][z_stride] Montgomery.
int r1cs_eval(const R1csDev& r, const void* z_mont, size_t z_stride_elems, size_t count, size_t m, void* abc, cudaStream_t st);

// h = witness_map: abc is consumed (a,b,c transformed in place with the two scratch buffers of the same size).
// out_h: [count][h_stride] canonical Fr.
int witness_map_run(const NttDomain& d, void* abc, void* scratch1, void* scratch2, size_t count, void* out_h,
                    size_t h_stride_elems, cudaStream_t st);

// canonical <-> Montgomery on arrays of Fr
int fr_to_mont(const void* in, void* out, size_t n, cudaStream_t st);
int fr_from_mont(const void* in, void* out, size_t n, cudaStream_t st);

}  // namespace mp
