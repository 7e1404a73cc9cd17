// Diagnostic entry points: element-wise field / group ops for the parity tests, and the integer-pipe
// microbenchmark that supplies the MSM roofline denominator.
#include "common.cuh"

namespace mp {

template <class F>
__global__ void k_field_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    constexpr int N = F::N;
    F x, y;
#pragma unroll
    for (int k = 0; k < N; k++) { x.l[k] = a[i * N + k]; y.l[k] = b ? b[i * N + k] : 0; }
    x = x.to_mont();
    y = y.to_mont();
    F r;
    switch (op) {
        case 0: r = x + y; break;
        case 1: r = x - y; break;
        case 2: r = x * y; break;
        case 3: r = x.sqr(); break;
        case 4: r = x.inv(); break;
        case 6: r = x.inv_gcd(); break;
        case 7: r = x.inv_kaliski(); break;
        default: r = x.neg(); break;
    }
    r = r.from_mont();
#pragma unroll
    for (int k = 0; k < N; k++) out[i * N + k] = r.l[k];
}

template <class F>
__global__ void k_group_op(int op, const uint32_t* a, const uint32_t* b, const uint32_t* k, uint32_t* out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    constexpr int W = Affine<F>::WORDS;
    Affine<F> pa = Affine<F>::load(a + i * W);
    XYZZ<F> r;
    if (op == 0) {
        Affine<F> pb = Affine<F>::load(b + i * W);
        r = XYZZ<F>::from_affine(pa).add_mixed_cold(pb);
    } else if (op == 1) {
        r = XYZZ<F>::dbl_affine(pa);
    } else {
        r = XYZZ<F>::inf();
        const uint32_t* s = k + i * 8;
        for (int bit = 254; bit >= 0; bit--) {
            r = r.dbl();
            if ((s[bit >> 5] >> (bit & 31)) & 1) r = r.add_mixed_cold(pa);
        }
    }
    r.to_affine().store(out + i * W);
}

// ---- integer-pipe microbenchmarks ----------------------------------------------------------------------
// 4 independent carry chains of 8 fused (mad.lo.cc, madc.hi.cc) pairs per thread: the IMAD.WIDE.U32(.X) shape the
// Montgomery multiply is made of.  `a` changes every iteration so nothing can be hoisted or strength-reduced.
__global__ void k_imad_wide(uint32_t* out, uint32_t x, uint32_t y, int iters) {
    uint32_t acc[4][16];
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
        for (int j = 0; j < 16; j++) acc[c][j] = threadIdx.x + j + c;
    uint32_t a = x + threadIdx.x, b = y + blockIdx.x;
    for (int it = 0; it < iters; it++) {
        a = a * 1664525u + 1013904223u;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            mad_wide_cc(acc[c][0], acc[c][1], a, b);
#pragma unroll
            for (int j = 2; j < 16; j += 2) madc_wide_cc(acc[c][j], acc[c][j + 1], a, b);
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
        for (int j = 0; j < 16; j++) s ^= acc[c][j];
    if (s == 0xdeadbeefu) out[0] = s;
}

// Production multiply throughput: 4 independent Fq product chains per thread.
__global__ void k_fq_mul_rate(uint32_t* out, int iters) {
    Fq a[2], b;
    for (int j = 0; j < 2; j++)
#pragma unroll
        for (int k = 0; k < 12; k++) a[j].l[k] = FQ_ONE[k] + threadIdx.x * (j + 1) + k;
#pragma unroll
    for (int k = 0; k < 12; k++) b.l[k] = FQ_R2[k] ^ blockIdx.x;
    b.l[11] &= 0x0fffffffu;
    for (int j = 0; j < 2; j++) a[j].l[11] &= 0x0fffffffu;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < 2; j++) a[j] = a[j] * b;
    }
    uint32_t s = 0;
    for (int j = 0; j < 2; j++)
#pragma unroll
        for (int k = 0; k < 12; k++) s ^= a[j].l[k];
    if (s == 0xdeadbeefu) out[0] = s;
}

}  // namespace mp

using namespace mp;

extern "C" {

int mp_debug_field_op(int device, int field, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
    if (!a || !out || (field != 0 && field != 1) || op < 0 || op > 7) return MP_ERR_INVALID_ARG;
    if ((op <= 2) && !b) return MP_ERR_INVALID_ARG;
    MP_TRY(use_device(device));
    if (n == 0) return MP_OK;
    size_t bytes = n * (field == 0 ? 48 : 32);
    DevBuf da, db, dout;
    MP_TRY(da.alloc(bytes));
    MP_TRY(dout.alloc(bytes));
    MP_CUDA_TRY(cudaMemcpy(da.p, a, bytes, cudaMemcpyHostToDevice));
    if (b) {
        MP_TRY(db.alloc(bytes));
        MP_CUDA_TRY(cudaMemcpy(db.p, b, bytes, cudaMemcpyHostToDevice));
    }
    if (field == 0)
        k_field_op<Fq><<<div_up(n, 128), 128>>>(op, da.as<uint32_t>(), b ? db.as<uint32_t>() : nullptr, dout.as<uint32_t>(), n);
    else
        k_field_op<Fr><<<div_up(n, 128), 128>>>(op, da.as<uint32_t>(), b ? db.as<uint32_t>() : nullptr, dout.as<uint32_t>(), n);
    MP_KERNEL_CHECK();
    MP_CUDA_TRY(cudaMemcpy(out, dout.p, bytes, cudaMemcpyDeviceToHost));
    return MP_OK;
}

int mp_debug_group_op(int device, int group, int op, const uint8_t* a, const uint8_t* b, const uint64_t* k,
                      uint8_t* out, size_t n) {
    if (!a || !out || (group != 1 && group != 2) || op < 0 || op > 2) return MP_ERR_INVALID_ARG;
    if ((op == 0 && !b) || (op == 2 && !k)) return MP_ERR_INVALID_ARG;
    MP_TRY(use_device(device));
    if (n == 0) return MP_OK;
    size_t pb = (group == 1 ? MP_G1_BYTES : MP_G2_BYTES);
    DevBuf da, db, dk, dout;
    MP_TRY(da.alloc(n * pb));
    MP_TRY(dout.alloc(n * pb));
    MP_CUDA_TRY(cudaMemcpy(da.p, a, n * pb, cudaMemcpyHostToDevice));
    if (b) {
        MP_TRY(db.alloc(n * pb));
        MP_CUDA_TRY(cudaMemcpy(db.p, b, n * pb, cudaMemcpyHostToDevice));
    }
    if (k) {
        MP_TRY(dk.alloc(n * 32));
        MP_CUDA_TRY(cudaMemcpy(dk.p, k, n * 32, cudaMemcpyHostToDevice));
    }
    if (group == 1) {
        MP_TRY(points_from_ark_g1(da.p, da.p, n, 0));
        if (b) MP_TRY(points_from_ark_g1(db.p, db.p, n, 0));
        k_group_op<Fq><<<div_up(n, 64), 64>>>(op, da.as<uint32_t>(), db.as<uint32_t>(), dk.as<uint32_t>(), dout.as<uint32_t>(), n);
        MP_KERNEL_CHECK();
        MP_TRY(points_to_ark_g1(dout.p, dout.p, n, 0));
    } else {
        MP_TRY(points_from_ark_g2(da.p, da.p, n, 0));
        if (b) MP_TRY(points_from_ark_g2(db.p, db.p, n, 0));
        k_group_op<Fq2><<<div_up(n, 64), 64>>>(op, da.as<uint32_t>(), db.as<uint32_t>(), dk.as<uint32_t>(), dout.as<uint32_t>(), n);
        MP_KERNEL_CHECK();
        MP_TRY(points_to_ark_g2(dout.p, dout.p, n, 0));
    }
    MP_CUDA_TRY(cudaMemcpy(out, dout.p, n * pb, cudaMemcpyDeviceToHost));
    return MP_OK;
}

int mp_debug_int_pipe_rate(int device, double* out_wide_mac_per_s, double* out_fq_mul_per_s) {
    MP_TRY(use_device(device));
    int sms = 0;
    MP_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    DevBuf sink;
    MP_TRY(sink.alloc(64));
    EventTimer timer;
    float ms = 0;
    {
        const int iters = 16384, threads = 512, blocks = sms * 2;
        k_imad_wide<<<blocks, threads>>>(sink.as<uint32_t>(), 3, 5, iters);  // warm-up (also ramps the clocks)
        double best = 0;
        for (int rep = 0; rep < 5; rep++) {
            MP_TRY(timer.start(0));
            k_imad_wide<<<blocks, threads>>>(sink.as<uint32_t>(), 3, 5, iters);
            MP_TRY(timer.stop(0));
            MP_TRY(timer.elapsed_ms(&ms));
            double rate = (double)blocks * threads * iters * 32.0 / (ms * 1e-3);
            if (rate > best) best = rate;
        }
        if (out_wide_mac_per_s) *out_wide_mac_per_s = best;
    }
    {
        const int iters = 8192, threads = 256, blocks = sms * 8;
        k_fq_mul_rate<<<blocks, threads>>>(sink.as<uint32_t>(), iters);
        double best = 0;
        for (int rep = 0; rep < 5; rep++) {
            MP_TRY(timer.start(0));
            k_fq_mul_rate<<<blocks, threads>>>(sink.as<uint32_t>(), iters);
            MP_TRY(timer.stop(0));
            MP_TRY(timer.elapsed_ms(&ms));
            double rate = (double)blocks * threads * iters * 2.0 / (ms * 1e-3);
            if (rate > best) best = rate;
        }
        if (out_fq_mul_per_s) *out_fq_mul_per_s = best;
    }
    MP_KERNEL_CHECK();
    return MP_OK;
}

}  // extern "C"
