// Integer-pipe microbenchmarks for sm_100a (run under gpurun).  Prints issue cost in SM cycles per
// warp-instruction per SMSP for several instruction shapes used by big-integer multiplication.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096

// 1. independent mad.wide.u32 (IMAD.WIDE.U32, no carry)
__global__ void k_wide(uint64_t* out, uint32_t x, uint32_t y) {
    uint64_t acc[8];
    for (int j = 0; j < 8; j++) acc[j] = threadIdx.x + j;
    uint32_t a = x + threadIdx.x, b = y + blockIdx.x;
    for (int it = 0; it < ITERS; it++) {
        a = a * 1664525u + 1013904223u;
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int j = 0; j < 8; j++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"(a + u), "r"(b ^ (j * 0x9e3779b9u)));
    }
    uint64_t s = 0;
    for (int j = 0; j < 8; j++) s ^= acc[j];
    if (s == 0x123456789abcdefull) out[0] = s;
}
// 2. carry chains: 4 independent chains of 8 fused (lo.cc, hi.cc) pairs -> IMAD.WIDE.U32.X
__global__ void k_wide_carry(uint32_t* out, uint32_t x, uint32_t y) {
    uint32_t acc[4][16];
    for (int c = 0; c < 4; c++) for (int j = 0; j < 16; j++) acc[c][j] = threadIdx.x + j + c;
    uint32_t a = x + threadIdx.x, b = y + blockIdx.x;
    for (int it = 0; it < ITERS; it++) {
        a = a * 1664525u + 1013904223u;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[c][0]), "+r"(acc[c][1]) : "r"(a), "r"(b));
#pragma unroll
            for (int j = 2; j < 16; j += 2)
                asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[c][j]), "+r"(acc[c][j + 1]) : "r"(a), "r"(b));
        }
    }
    uint32_t s = 0;
    for (int c = 0; c < 4; c++) for (int j = 0; j < 16; j++) s ^= acc[c][j];
    if (s == 0xdeadbeefu) out[0] = s;
}
// 3. separate IMAD lo and IMAD.HI (no carry, 32-bit results)
__global__ void k_lo_hi(uint32_t* out, uint32_t x, uint32_t y) {
    uint32_t acc[16];
    for (int j = 0; j < 16; j++) acc[j] = threadIdx.x + j;
    uint32_t a = x + threadIdx.x, b = y + blockIdx.x;
    for (int it = 0; it < ITERS; it++) {
        a = a * 1664525u + 1013904223u;
#pragma unroll
        for (int u = 0; u < 2; u++)
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[j]) : "r"(a + u), "r"(b ^ (j * 0x9e3779b9u)));
                asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(acc[j + 1]) : "r"(a + u), "r"(b ^ (j * 0x9e3779b9u)));
            }
    }
    uint32_t s = 0;
    for (int j = 0; j < 16; j++) s ^= acc[j];
    if (s == 0xdeadbeefu) out[0] = s;
}
// 4. mad.wide + independent IADD3 (does the ALU pipe co-issue?)  32 wide + 32 add per iteration
__global__ void k_wide_plus_add(uint64_t* out, uint32_t x, uint32_t y) {
    uint64_t acc[8];
    uint32_t t[8];
    for (int j = 0; j < 8; j++) { acc[j] = threadIdx.x + j; t[j] = j; }
    uint32_t a = x + threadIdx.x, b = y + blockIdx.x;
    for (int it = 0; it < ITERS; it++) {
        a = a * 1664525u + 1013904223u;
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int j = 0; j < 8; j++) {
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"(a + u), "r"(b ^ (j * 0x9e3779b9u)));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(t[j]) : "r"(a));
            }
    }
    uint64_t s = 0;
    for (int j = 0; j < 8; j++) s ^= acc[j] + t[j];
    if (s == 0x123456789abcdefull) out[0] = s;
}
// 5. DFMA
__global__ void k_dfma(double* out, double x, double y) {
    double acc[8];
    for (int j = 0; j < 8; j++) acc[j] = threadIdx.x + j;
    double a = x + threadIdx.x, b = y + blockIdx.x;
    for (int it = 0; it < ITERS; it++) {
        a = a * 1.0000001 + 0.5;
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int j = 0; j < 8; j++) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(acc[j]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int j = 0; j < 8; j++) s += acc[j];
    if (s == 1.2345) out[0] = s;
}
// 6. carry chain of plain adds (IADD3.X) : 4 chains x 8
__global__ void k_addc(uint32_t* out, uint32_t x) {
    uint32_t acc[4][8];
    for (int c = 0; c < 4; c++) for (int j = 0; j < 8; j++) acc[c][j] = threadIdx.x + j + c;
    uint32_t a = x + threadIdx.x;
    for (int it = 0; it < ITERS; it++) {
        a = a * 1664525u + 1013904223u;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(acc[c][0]) : "r"(a));
#pragma unroll
            for (int j = 1; j < 8; j++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(acc[c][j]) : "r"(a));
        }
    }
    uint32_t s = 0;
    for (int c = 0; c < 4; c++) for (int j = 0; j < 8; j++) s ^= acc[c][j];
    if (s == 0xdeadbeefu) out[0] = s;
}
// 7. wide MAC with carry-out only at chain end: pairs (wide no-carry) interleaved with carry chains 1:1
__global__ void k_wide_mix(uint32_t* out, uint32_t x, uint32_t y) {
    uint32_t acc[2][16];
    uint64_t w[8];
    for (int c = 0; c < 2; c++) for (int j = 0; j < 16; j++) acc[c][j] = threadIdx.x + j + c;
    for (int j = 0; j < 8; j++) w[j] = j + threadIdx.x;
    uint32_t a = x + threadIdx.x, b = y + blockIdx.x;
    for (int it = 0; it < ITERS; it++) {
        a = a * 1664525u + 1013904223u;
#pragma unroll
        for (int c = 0; c < 2; c++) {
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[c][0]), "+r"(acc[c][1]) : "r"(a), "r"(b));
#pragma unroll
            for (int j = 2; j < 16; j += 2)
                asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[c][j]), "+r"(acc[c][j + 1]) : "r"(a), "r"(b));
#pragma unroll
            for (int j = 0; j < 8; j++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[j]) : "r"(a + c), "r"(b ^ (j * 0x9e3779b9u)));
        }
    }
    uint32_t s = 0;
    for (int c = 0; c < 2; c++) for (int j = 0; j < 16; j++) s ^= acc[c][j];
    for (int j = 0; j < 8; j++) s ^= (uint32_t)w[j];
    if (s == 0xdeadbeefu) out[0] = s;
}

template <class K, class... A>
static void run(const char* name, double instr_per_thread, int threads, int blocks_per_sm, K kern, A... args) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<sms * blocks_per_sm, threads>>>(args...);
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        kern<<<sms * blocks_per_sm, threads>>>(args...);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double warps_per_smsp = (double)threads * blocks_per_sm / 32 / 4;
    double warp_instr_per_smsp = instr_per_thread * warps_per_smsp;
    double cycles = best * 1e-3 * clk * 1e3;  // at max clock (upper bound on cycles)
    printf("%-28s %8.3f ms  %.3g thread-instr/s  ~%.2f clk/warp-instr/SMSP (at %d MHz)\n", name, best,
           instr_per_thread * threads * blocks_per_sm * sms / (best * 1e-3), cycles / warp_instr_per_smsp, clk / 1000);
}

int main() {
    void* sink; cudaMalloc(&sink, 256);
    run("mad.wide (no carry)", ITERS * 32.0, 512, 4, k_wide, (uint64_t*)sink, 3u, 5u);
    run("lo.cc/hi.cc chains (.X)", ITERS * 32.0, 512, 2, k_wide_carry, (uint32_t*)sink, 3u, 5u);
    run("mad.lo + mad.hi separate", ITERS * 32.0, 512, 4, k_lo_hi, (uint32_t*)sink, 3u, 5u);
    run("mad.wide + add (64 instr)", ITERS * 64.0, 512, 4, k_wide_plus_add, (uint64_t*)sink, 3u, 5u);
    run("dfma", ITERS * 32.0, 512, 4, k_dfma, (double*)sink, 3.0, 5.0);
    run("add.cc chains", ITERS * 32.0, 512, 4, k_addc, (uint32_t*)sink, 3u);
    run("carry(16)+wide(16) mix", ITERS * 32.0, 512, 2, k_wide_mix, (uint32_t*)sink, 3u, 5u);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
