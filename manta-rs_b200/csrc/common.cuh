// Host-side plumbing shared by the translation units of libmantaprover.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/mantaprover.h"
#include "ec.cuh"

namespace mp {

// thread-local failure text behind mp_last_error_detail()
void set_error_detail(const char* fmt, ...);

#define MP_CUDA_TRY(expr)                                                                       \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            mp::set_error_detail("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return (_e == cudaErrorMemoryAllocation) ? MP_ERR_OOM                               \
                   : (_e == cudaErrorNoDevice || _e == cudaErrorInsufficientDriver || _e == cudaErrorInvalidDevice) ? MP_ERR_NO_DEVICE \
                                                                                                : MP_ERR_CUDA; \
        }                                                                                       \
    } while (0)

#define MP_TRY(expr)               \
    do {                           \
        int _rc = (expr);          \
        if (_rc != MP_OK) return _rc; \
    } while (0)

// every kernel launch of the library is followed by this check; it also feeds the per-batch launch count
uint64_t& kernel_launch_counter();
#define MP_KERNEL_CHECK()                 \
    do {                                  \
        mp::kernel_launch_counter()++;    \
        MP_CUDA_TRY(cudaGetLastError());  \
    } while (0)

// Selects the device and verifies it is an sm_100-class part (no fallback path exists).
int use_device(int device);

// bytes handed out by DevBuf::alloc on this thread (batch objects report the sum of their buffers from it)
uint64_t& dev_alloc_counter();

// RAII device buffer (freed on scope exit; never throws)
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    int alloc(size_t n) {
        release();
        if (n == 0) n = 16;
        MP_CUDA_TRY(cudaMalloc(&p, n));
        bytes = n;
        dev_alloc_counter() += n;
        return MP_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// RAII pair of timing events around a region of one stream (destroyed on every return path; never throws)
struct EventTimer {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    EventTimer() = default;
    EventTimer(const EventTimer&) = delete;
    EventTimer& operator=(const EventTimer&) = delete;
    ~EventTimer() {
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
    }
    int start(cudaStream_t st) {
        if (!e0) MP_CUDA_TRY(cudaEventCreate(&e0));
        if (!e1) MP_CUDA_TRY(cudaEventCreate(&e1));
        MP_CUDA_TRY(cudaEventRecord(e0, st));
        return MP_OK;
    }
    int stop(cudaStream_t st) {
        MP_CUDA_TRY(cudaEventRecord(e1, st));
        return MP_OK;
    }
    int elapsed_ms(float* ms) {  // waits for the stop event
        MP_CUDA_TRY(cudaEventSynchronize(e1));
        MP_CUDA_TRY(cudaEventElapsedTime(ms, e0, e1));
        return MP_OK;
    }
};

static inline unsigned div_up(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// ---- ark-serialize <-> device layout ----------------------------------------------------------------
// Uncompressed ark points and device `Affine<F>` have the same size (G1 96 B, G2 192 B); the kernels
// below convert in place or out of place between canonical little-endian bytes (infinity flag 0x40 in
// the top byte of y) and Montgomery limbs with (0,0) = infinity.
template <class F> __global__ void k_points_from_ark(const uint32_t* in, uint32_t* out, size_t n);
template <class F> __global__ void k_points_to_ark(const uint32_t* in, uint32_t* out, size_t n);

int points_from_ark_g1(const void* d_in, void* d_out, size_t n, cudaStream_t st);
int points_from_ark_g2(const void* d_in, void* d_out, size_t n, cudaStream_t st);
int points_to_ark_g1(const void* d_in, void* d_out, size_t n, cudaStream_t st);
int points_to_ark_g2(const void* d_in, void* d_out, size_t n, cudaStream_t st);

}  // namespace mp
