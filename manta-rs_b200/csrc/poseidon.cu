// Batched Poseidon permutation over BLS12-381 Fr (witness-side Fr work next to the prover).
//
// Replaces, for batches of independent states, `Permutation::permute` of manta-pay
// (manta-pay/src/crypto/poseidon/mod.rs:385-421,515-518): round = add round keys, S-box x^5 on every element (full
// round) or on element 0 (partial round), MDS multiply; FULL_ROUNDS / 2 full rounds, PARTIAL_ROUNDS partial rounds,
// FULL_ROUNDS / 2 full rounds.  Parameters are the caller's (round keys in round order, MDS row-major), canonical in,
// Montgomery on the device.  One thread per state, the state in registers; pinned by the reference's own known-answer
// vector permutation_hardcoded_test/width3 (hash.rs:248-258) in tests/test_gpu_parity.py.
#include "common.cuh"

namespace mp {

constexpr int POSEIDON_MAX_WIDTH = 8;

template <int W>
__global__ void __launch_bounds__(128) k_poseidon(const uint32_t* __restrict__ rk, const uint32_t* __restrict__ mds, uint32_t* states,
                                                  size_t count, int half_full, int partial) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= count) return;
    Fr st[W];
#pragma unroll
    for (int j = 0; j < W; j++) st[j] = Fr::load(states + (i * W + j) * 8).to_mont();
    const int rounds = 2 * half_full + partial;
    for (int r = 0; r < rounds; r++) {
        const bool full = r < half_full || r >= half_full + partial;
#pragma unroll
        for (int j = 0; j < W; j++) st[j] = st[j] + Fr::load_ro(rk + ((size_t)r * W + j) * 8);
#pragma unroll
        for (int j = 0; j < W; j++) {
            if (j == 0 || full) {
                Fr x2 = st[j].sqr();
                st[j] = x2.sqr() * st[j];
            }
        }
        Fr nx[W];
#pragma unroll
        for (int a = 0; a < W; a++) {
            Fr acc = Fr::load_ro(mds + (size_t)(a * W) * 8) * st[0];
#pragma unroll
            for (int b = 1; b < W; b++) acc = acc + Fr::load_ro(mds + (size_t)(a * W + b) * 8) * st[b];
            nx[a] = acc;
        }
#pragma unroll
        for (int j = 0; j < W; j++) st[j] = nx[j];
    }
#pragma unroll
    for (int j = 0; j < W; j++) st[j].from_mont().store(states + (i * W + j) * 8);
}

__global__ void k_fr_to_mont_inplace(uint32_t* v, size_t n) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) Fr::load(v + i * 8).to_mont().store(v + i * 8);
}

template <int W>
static void launch_poseidon(const uint32_t* rk, const uint32_t* mds, uint32_t* st, size_t count, int hf, int p) {
    k_poseidon<W><<<div_up(count, 128), 128>>>(rk, mds, st, count, hf, p);
}

}  // namespace mp

using namespace mp;

extern "C" int mp_poseidon_permute(int device, int width, int full_rounds, int partial_rounds, const uint64_t* round_keys,
                                   const uint64_t* mds, uint64_t* states, size_t count, float* out_device_ms) {
    if (width < 2 || width > POSEIDON_MAX_WIDTH || full_rounds < 2 || (full_rounds & 1) || partial_rounds < 0 || !round_keys || !mds ||
        (count && !states))
        return MP_ERR_INVALID_ARG;
    MP_TRY(use_device(device));
    if (out_device_ms) *out_device_ms = 0;
    if (count == 0) return MP_OK;
    const size_t n_rk = (size_t)(full_rounds + partial_rounds) * width, n_mds = (size_t)width * width;
    DevBuf d_par, d_st;
    MP_TRY(d_par.alloc((n_rk + n_mds) * 32));
    MP_TRY(d_st.alloc(count * width * 32));
    MP_CUDA_TRY(cudaMemcpy(d_par.p, round_keys, n_rk * 32, cudaMemcpyHostToDevice));
    MP_CUDA_TRY(cudaMemcpy(d_par.as<char>() + n_rk * 32, mds, n_mds * 32, cudaMemcpyHostToDevice));
    MP_CUDA_TRY(cudaMemcpy(d_st.p, states, count * width * 32, cudaMemcpyHostToDevice));
    k_fr_to_mont_inplace<<<div_up(n_rk + n_mds, 128), 128>>>(d_par.as<uint32_t>(), n_rk + n_mds);
    MP_KERNEL_CHECK();
    EventTimer timer;
    MP_TRY(timer.start(0));
    const uint32_t* rk = d_par.as<uint32_t>();
    const uint32_t* md = rk + n_rk * 8;
    uint32_t* st = d_st.as<uint32_t>();
    const int hf = full_rounds / 2;
    switch (width) {
        case 2: launch_poseidon<2>(rk, md, st, count, hf, partial_rounds); break;
        case 3: launch_poseidon<3>(rk, md, st, count, hf, partial_rounds); break;
        case 4: launch_poseidon<4>(rk, md, st, count, hf, partial_rounds); break;
        case 5: launch_poseidon<5>(rk, md, st, count, hf, partial_rounds); break;
        case 6: launch_poseidon<6>(rk, md, st, count, hf, partial_rounds); break;
        case 7: launch_poseidon<7>(rk, md, st, count, hf, partial_rounds); break;
        default: launch_poseidon<8>(rk, md, st, count, hf, partial_rounds); break;
    }
    MP_KERNEL_CHECK();
    MP_TRY(timer.stop(0));
    MP_CUDA_TRY(cudaMemcpy(states, d_st.p, count * width * 32, cudaMemcpyDeviceToHost));
    float ms = 0;
    MP_TRY(timer.elapsed_ms(&ms));
    if (out_device_ms) *out_device_ms = ms;
    return MP_OK;
}
