// Bucket-accumulation inner loop microbenchmark: G1 / G2 XYZZ mixed adds per second for one arithmetic variant
// (selected with -D flags understood by fp.cuh / ec.cuh).  Build one binary per variant; run under gpurun.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../manta-rs_b200/csrc/ec.cuh"
using namespace mp;

template <class F>
__global__ void __launch_bounds__(128) k_fill(uint32_t* tab, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // distinct valid points: multiples of the generator by repeated doubling/adding (cheap: P_i = 2*P_{i-1} chain per thread)
    Affine<F> g;
    if (FieldWords<F>::W == 12) g = Affine<F>::load(&G1_GEN[0][0]); else g = Affine<F>::load(&G2_GEN[0][0]);
    XYZZ<F> p = XYZZ<F>::from_affine(g);
    int k = i + 2;
    XYZZ<F> r = XYZZ<F>::inf();
    for (int b = 20; b >= 0; b--) { r = r.dbl(); if ((k >> b) & 1) r = r.add_mixed_cold(g); }
    r.to_affine().store(tab + (size_t)i * Affine<F>::WORDS);
}

template <class F>
__global__ void __launch_bounds__(128) k_madd(const uint32_t* __restrict__ tab, int n, int len, uint32_t* out) {
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t x = tid * 2654435761u + 12345u;
    XYZZ<F> acc = XYZZ<F>::inf();
    for (int e = 0; e < len; e++) {
        x = x * 1664525u + 1013904223u;
        uint32_t idx = (x >> 8) % (uint32_t)n;
        Affine<F> p = Affine<F>::load_ro(tab + (size_t)idx * Affine<F>::WORDS);
        if (x & 1) p.y = p.y.neg();
        acc = acc.add_mixed(p);
    }
    acc.store(out + (size_t)tid * XYZZ<F>::WORDS);
}

template <class F>
static void run(const char* name, int blocks_per_sm_hint) {
    const int n = 1 << 15, len = 24;
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t *tab, *out;
    cudaMalloc(&tab, (size_t)n * Affine<F>::WORDS * 4);
    const int threads = 128, blocks = sms * 64;
    cudaMalloc(&out, (size_t)blocks * threads * XYZZ<F>::WORDS * 4);
    k_fill<F><<<(n + 127) / 128, 128>>>(tab, n);
    k_madd<F><<<blocks, threads>>>(tab, n, len, out);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0); k_madd<F><<<blocks, threads>>>(tab, n, len, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k_madd<F>);
    int occ; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_madd<F>, threads, 0);
    // checksum
    uint32_t h[8]; cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost);
    printf("%-4s %-40s regs=%3d local=%4zu occ=%d blk/SM  %8.3f ms  %.3f G madd/s  (chk %08x)  %s\n", name, VARIANT, fa.numRegs, fa.localSizeBytes, occ, best,
           (double)blocks * threads * (len - 1) / (best * 1e-3) / 1e9, h[0] ^ h[5], cudaGetErrorString(cudaGetLastError()));
    cudaFree(tab); cudaFree(out);
}

int main() {
    run<Fq>("G1", 3);
    run<Fq2>("G2", 2);
    return 0;
}
