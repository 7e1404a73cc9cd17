// Library basics: error reporting, device selection, ark-serialize <-> device point layout.
#include <cstdarg>

#include "common.cuh"

namespace mp {

static thread_local char g_detail[512] = "";

uint64_t& kernel_launch_counter() {
    static thread_local uint64_t n = 0;
    return n;
}

uint64_t& dev_alloc_counter() {
    static thread_local uint64_t n = 0;
    return n;
}

void set_error_detail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_detail, sizeof(g_detail), fmt, ap);
    va_end(ap);
}

int use_device(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error_detail("no CUDA device: %s", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return MP_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= count) {
        set_error_detail("device %d out of range (have %d)", device, count);
        return MP_ERR_INVALID_ARG;
    }
    MP_CUDA_TRY(cudaSetDevice(device));
    int major = 0;
    MP_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    if (major != 10) {
        set_error_detail("device %d is sm_%d0-class; this library only carries sm_100a code", device, major);
        return MP_ERR_NO_DEVICE;
    }
    return MP_OK;
}

// ---- point conversion kernels ---------------------------------------------------------------------------
template <class F> struct Coord;
template <> struct Coord<Fq> {
    static constexpr int W = 12;
    MP_DEV static Fq load_canon(const uint32_t* p, bool strip) {
        Fq v;
#pragma unroll
        for (int i = 0; i < 12; i++) v.l[i] = p[i];
        if (strip) v.l[11] &= 0x3fffffffu;
        return v.to_mont();
    }
    MP_DEV static void store_canon(uint32_t* p, const Fq& v) {
        Fq c = v.from_mont();
#pragma unroll
        for (int i = 0; i < 12; i++) p[i] = c.l[i];
    }
};
template <> struct Coord<Fq2> {
    static constexpr int W = 24;
    MP_DEV static Fq2 load_canon(const uint32_t* p, bool strip) {
        return {Coord<Fq>::load_canon(p, false), Coord<Fq>::load_canon(p + 12, strip)};
    }
    MP_DEV static void store_canon(uint32_t* p, const Fq2& v) {
        Coord<Fq>::store_canon(p, v.c0);
        Coord<Fq>::store_canon(p + 12, v.c1);
    }
};

template <class F>
__global__ void k_points_from_ark(const uint32_t* in, uint32_t* out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    constexpr int W = Coord<F>::W;
    const uint32_t* p = in + i * 2 * W;
    bool inf = (p[2 * W - 1] >> 30) & 1;  // 0x40 of the last byte
    Affine<F> a;
    if (inf) {
        a = Affine<F>::inf();
    } else {
        a.x = Coord<F>::load_canon(p, false);
        a.y = Coord<F>::load_canon(p + W, true);
    }
    uint32_t* q = out + i * 2 * W;
    a.store(q);
}

template <class F>
__global__ void k_points_to_ark(const uint32_t* in, uint32_t* out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    constexpr int W = Coord<F>::W;
    Affine<F> a = Affine<F>::load(in + i * 2 * W);
    uint32_t* q = out + i * 2 * W;
    if (a.is_inf()) {
        for (int k = 0; k < 2 * W; k++) q[k] = 0;
        q[2 * W - 1] = 0x40000000u;
    } else {
        Coord<F>::store_canon(q, a.x);
        Coord<F>::store_canon(q + W, a.y);
    }
}

int points_from_ark_g1(const void* in, void* out, size_t n, cudaStream_t st) {
    if (n == 0) return MP_OK;
    k_points_from_ark<Fq><<<div_up(n, 128), 128, 0, st>>>((const uint32_t*)in, (uint32_t*)out, n);
    MP_KERNEL_CHECK();
    return MP_OK;
}
int points_from_ark_g2(const void* in, void* out, size_t n, cudaStream_t st) {
    if (n == 0) return MP_OK;
    k_points_from_ark<Fq2><<<div_up(n, 128), 128, 0, st>>>((const uint32_t*)in, (uint32_t*)out, n);
    MP_KERNEL_CHECK();
    return MP_OK;
}
int points_to_ark_g1(const void* in, void* out, size_t n, cudaStream_t st) {
    if (n == 0) return MP_OK;
    k_points_to_ark<Fq><<<div_up(n, 128), 128, 0, st>>>((const uint32_t*)in, (uint32_t*)out, n);
    MP_KERNEL_CHECK();
    return MP_OK;
}
int points_to_ark_g2(const void* in, void* out, size_t n, cudaStream_t st) {
    if (n == 0) return MP_OK;
    k_points_to_ark<Fq2><<<div_up(n, 128), 128, 0, st>>>((const uint32_t*)in, (uint32_t*)out, n);
    MP_KERNEL_CHECK();
    return MP_OK;
}

}  // namespace mp

extern "C" {

const char* mp_strerror(int code) {
    switch (code) {
        case MP_OK: return "ok";
        case MP_ERR_INVALID_ARG: return "invalid argument";
        case MP_ERR_CUDA: return "CUDA runtime error";
        case MP_ERR_NO_DEVICE: return "no usable sm_100 CUDA device (this library has no CPU fallback)";
        case MP_ERR_OOM: return "out of device memory";
        case MP_ERR_FORMAT: return "malformed serialized input";
        case MP_ERR_UNSUPPORTED: return "unsupported size or option";
        default: return "unknown error";
    }
}

const char* mp_last_error_detail(void) { return mp::g_detail; }

int mp_device_count(int* out_count) {
    if (!out_count) return MP_ERR_INVALID_ARG;
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) {
        mp::set_error_detail("cudaGetDeviceCount: %s", cudaGetErrorString(e));
        *out_count = 0;
        return MP_ERR_NO_DEVICE;
    }
    *out_count = c;
    return MP_OK;
}

}  // extern "C"
