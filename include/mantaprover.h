/*
 * mantaprover.h — C ABI of the B200-native Groth16 proving backend (BLS12-381).
 *
 * This is the boundary a Rust shim binds to replace the body of
 *   manta-crypto/src/arkworks/groth16.rs:588-600   (`ProofSystem::prove` for `Groth16<E>`)
 * i.e. the call `ArkGroth16::prove(&context.proving_key, compiler, &mut SizedRng(rng))` at :597, which in
 * upstream ark-groth16 0.3 is `create_random_proof` -> `create_proof`.
 * The shim keeps the trait signature (`manta-crypto/src/constraint.rs:87-95`), draws r and s from the caller's
 * rng exactly as `create_random_proof` does, finalizes the R1CS and hands the full assignment here; any non-zero
 * return code is collapsed into the unit `Error` of groth16.rs:50-60.  See INTEGRATION.md for the binding.
 *
 * Conventions
 *   - every function returns an int: 0 = MP_OK, otherwise an MP_ERR_* code; nothing throws or aborts;
 *   - the caller owns every host buffer for the duration of the call only;
 *   - scalars (Fr) are 4 little-endian uint64 limbs in CANONICAL (non-Montgomery) form, i.e. ark `into_repr()`;
 *   - points cross the boundary in ark-serialize 0.3 canonical byte form:
 *       uncompressed G1 = x(48 LE) | y(48 LE), G2 = x.c0 | x.c1 | y.c0 | y.c1, infinity flag 0x40 in the last byte;
 *       compressed   G1 = x(48 LE) with 0x80 = "y is the larger of {y,-y}", 0x40 = infinity in the last byte;
 *   - a context is bound to one CUDA device (one process per GPU); calls on one context are serialized
 *     internally, distinct contexts are independent.  There is NO CPU fallback: without a usable sm_100 device
 *     every entry point returns MP_ERR_NO_DEVICE / MP_ERR_CUDA.
 */
#ifndef MANTAPROVER_H
#define MANTAPROVER_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MP_API __attribute__((visibility("default")))
#else
#define MP_API
#endif

#define MP_OK 0
#define MP_ERR_INVALID_ARG 1
#define MP_ERR_CUDA 2
#define MP_ERR_NO_DEVICE 3
#define MP_ERR_OOM 4
#define MP_ERR_FORMAT 5
#define MP_ERR_UNSUPPORTED 6

#define MP_FR_LIMBS 4      /* uint64 limbs per Fr scalar */
#define MP_G1_BYTES 96     /* uncompressed */
#define MP_G2_BYTES 192    /* uncompressed */
#define MP_PROOF_BYTES 192 /* compressed A(48) | B(96) | C(48): `proof_as_bytes`, groth16.rs:184-195 */
#define MP_MAX_BATCH 21845 /* proofs per batch object (3 vectors per proof ride in gridDim.y of the NTT launches) */

typedef struct mp_ctx mp_ctx;
typedef struct mp_batch mp_batch;

/* ark_groth16::ProvingKey<Bls12_381> as owned by `ProvingContext<E>` (groth16.rs:208-245); every pointer
 * addresses ark uncompressed points.  `mp_pk_parse` fills this from the on-disk `ProvingContext` encoding. */
typedef struct mp_pk_view {
    const uint8_t* alpha_g1;  /* vk.alpha_g1            1 x G1 */
    const uint8_t* beta_g2;   /* vk.beta_g2             1 x G2 */
    const uint8_t* gamma_g2;  /* vk.gamma_g2            1 x G2 (unused by prove) */
    const uint8_t* delta_g2;  /* vk.delta_g2            1 x G2 */
    const uint8_t* gamma_abc_g1; uint64_t gamma_abc_len; /* vk.gamma_abc_g1 (unused by prove) */
    const uint8_t* beta_g1;   /* 1 x G1 */
    const uint8_t* delta_g1;  /* 1 x G1 */
    const uint8_t* a_query;    uint64_t a_len;    /* n x G1 */
    const uint8_t* b_g1_query; uint64_t b_g1_len; /* n x G1 */
    const uint8_t* b_g2_query; uint64_t b_g2_len; /* n x G2 */
    const uint8_t* h_query;    uint64_t h_len;    /* m-1 (ark generator) or m (MPC keys) x G1 */
    const uint8_t* l_query;    uint64_t l_len;    /* w x G1 */
} mp_pk_view;

/* Sparse R1CS matrices as produced by ark-relations `to_matrices()` after `finalize()`:
 * CSR per matrix; column index = instance variables first (0 is the constant 1) then witnesses;
 * coefficients canonical Fr, 4 x uint64 LE each. */
typedef struct mp_r1cs_view {
    uint64_t num_instance;    /* p, including the constant 1 */
    uint64_t num_witness;     /* w */
    uint64_t num_constraints; /* K */
    const uint64_t* a_row_ptr; const uint32_t* a_col; const uint64_t* a_coeff;
    const uint64_t* b_row_ptr; const uint32_t* b_col; const uint64_t* b_coeff;
    const uint64_t* c_row_ptr; const uint32_t* c_col; const uint64_t* c_coeff;
} mp_r1cs_view;

/* ---- library ------------------------------------------------------------------------------------------- */
MP_API const char* mp_strerror(int code);
MP_API const char* mp_last_error_detail(void); /* thread-local text of the last failure (CUDA error string etc.) */
MP_API int mp_device_count(int* out_count);

/* ---- proving-key file format (groth16.rs:268-303: `serialize_unchecked`, i.e. uncompressed, field order
 *      vk{alpha_g1,beta_g2,gamma_g2,delta_g2,gamma_abc_g1} beta_g1 delta_g1 a b_g1 b_g2 h l, u64-LE lengths) --- */
MP_API int mp_pk_parse(const uint8_t* data, size_t len, mp_pk_view* out);

/* ---- context: proving key + circuit matrices resident on one device ---------------------------------------
 * Replaces holding `&ProvingContext<E>` (groth16.rs:208-245).  Uploads the key, converts it to Montgomery
 * form and builds the per-window base tables on the device. */
MP_API int mp_ctx_create(const mp_pk_view* pk, const mp_r1cs_view* r1cs, int device, mp_ctx** out);
MP_API void mp_ctx_destroy(mp_ctx* ctx);
MP_API int mp_ctx_info(const mp_ctx* ctx, uint64_t* n_vars, uint64_t* n_instance, uint64_t* domain_size, uint64_t* device_bytes);

/* ---- prove: replaces ark_groth16::create_proof(circuit, pk, r, s) behind groth16.rs:597 ---------------------
 * z = full assignment [1, instance.., witness..] (n x 4 limbs), r/s = the two Fr draws of create_random_proof. */
MP_API int mp_prove(mp_ctx* ctx, const uint64_t* z, const uint64_t r[4], const uint64_t s[4], uint8_t out_proof[MP_PROOF_BYTES]);
/* The same proof when the host has already evaluated the constraint system (the first step of ark's `witness_map`): a, b, c
 * are the domain_size evaluations <A_i, z>, <B_i, z>, <C_i, z> (canonical, zero padded; a carries z_j at rows K + j, the
 * "dummy input constraints"), so the context's matrices are not consulted. */
MP_API int mp_prove_from_abc(mp_ctx* ctx, const uint64_t* z, const uint64_t* a, const uint64_t* b, const uint64_t* c, const uint64_t r[4],
                             const uint64_t s[4], uint8_t out_proof[MP_PROOF_BYTES]);
/* count independent proofs against the same context; z is count x n x 4 limbs contiguous, r/s count x 4. */
MP_API int mp_prove_batch(mp_ctx* ctx, size_t count, const uint64_t* z, const uint64_t* r, const uint64_t* s,
                   uint8_t* out_proofs /* count x 192 */);

/* Staged form of mp_prove_batch (what it does internally), so a harness can time the device-resident part:
 * upload = H2D of assignments, run = all kernels, download = D2H of proof bytes. */
MP_API int mp_batch_create(mp_ctx* ctx, size_t capacity, mp_batch** out);
/* high_priority = 1 puts the batch's streams at the greatest CUDA stream priority: with two batches in flight the
 * low-priority one only fills the gaps (latency-bound tails, copies) of the high-priority one. */
MP_API int mp_batch_create_ex(mp_ctx* ctx, size_t capacity, int high_priority, mp_batch** out);
MP_API void mp_batch_destroy(mp_batch* b);
MP_API int mp_batch_upload(mp_batch* b, size_t count, const uint64_t* z, const uint64_t* r, const uint64_t* s);
MP_API int mp_batch_run(mp_batch* b, float* out_device_ms /* nullable: CUDA-event time of the whole run */);
MP_API int mp_batch_download(mp_batch* b, uint8_t* out_proofs);
/* Asynchronous forms, for keeping two batches in flight so that the latency-bound tail of one (bucket-reduction level 2,
 * finishing kernel) and its host copies hide behind the kernels of the next:
 *   mp_batch_run_async  = mp_batch_run without the final synchronisation (inputs already uploaded);
 *   mp_batch_submit     = H2D of the inputs + all kernels + D2H of the proofs into out_proofs, all enqueued; host buffers
 *                         must stay valid (pinned memory makes the copies truly asynchronous) until mp_batch_wait;
 *   mp_batch_wait       = block until the batch is done; reports the CUDA-event time of its kernels. */
MP_API int mp_batch_run_async(mp_batch* b);
MP_API int mp_batch_submit(mp_batch* b, size_t count, const uint64_t* z, const uint64_t* r, const uint64_t* s, uint8_t* out_proofs);
MP_API int mp_batch_wait(mp_batch* b, float* out_device_ms);
/* per-phase CUDA-event times of the last mp_batch_run, in ms; names via mp_phase_name(i). Returns count. */
MP_API int mp_batch_phase_ms(const mp_batch* b, float* out_ms, int max_phases);
MP_API const char* mp_phase_name(int i);
/* The dominant kernel of the last mp_batch_run (round-1 k_ba_bwd<Fq>: the first tree level of the four G1 bucket
 * accumulations): its CUDA-event time and the number of affine additions it performed (5 Fq multiplications each). */
MP_API int mp_batch_dominant_kernel(mp_batch* b, float* out_ms, uint64_t* out_additions);
MP_API uint64_t mp_batch_kernel_launches(const mp_batch* b); /* kernels launched by the last mp_batch_run */
MP_API uint64_t mp_batch_device_bytes(const mp_batch* b);    /* device memory held by the batch object (all of it is allocated at creation) */
/* overlap = 1 (default): the latency-bound tail of the G2 reduction (for batches of <= 16 proofs the whole G2 MSM) runs on a
 * second stream next to the witness map and the G1 MSMs; overlap = 0:
 * every kernel on one stream in program order, so the per-phase CUDA-event times are those of the kernels alone. */
MP_API int mp_batch_set_overlap(mp_batch* b, int overlap);

/* ---- stand-alone kernels of the path (ark-ec `VariableBaseMSM::multi_scalar_mul`, ark-poly radix-2 domain;
 *      direct reference call sites: manta-benchmark/src/ecc.rs:62-118, manta-trusted-setup/src/groth16/mpc.rs:367-381) */
MP_API int mp_msm_g1(int device, const uint8_t* bases /* n x 96 */, const uint64_t* scalars /* n x 4 */, size_t n,
              uint8_t out_point[MP_G1_BYTES], float* out_device_ms);
MP_API int mp_msm_g2(int device, const uint8_t* bases /* n x 192 */, const uint64_t* scalars, size_t n,
              uint8_t out_point[MP_G2_BYTES], float* out_device_ms);
/* The same MSM against bases that stay resident on the device (proving keys and ceremony powers are fixed; only the
 * scalars change per call): create uploads the bases, converts them to Montgomery form and precomputes the window rows
 * 2^(c t) P_i that fit MP_MSM_TABLE_LIMIT_MB (default 8192) of device memory, run moves n <= n_bases scalars (the rest count
 * as zero, ark's `size = min(bases, scalars)`) and returns the uncompressed result.  group: 1 = G1, 2 = G2. */
typedef struct mp_msm_bases mp_msm_bases;
MP_API int mp_msm_bases_create(int device, int group, const uint8_t* bases, size_t n, mp_msm_bases** out);
MP_API int mp_msm_bases_run(mp_msm_bases* h, const uint64_t* scalars, size_t n, uint8_t* out_point, float* out_device_ms);
MP_API void mp_msm_bases_destroy(mp_msm_bases* h);
/* Sum of n affine points (ark uncompressed in, ark uncompressed out).  Combines the per-GPU partial results of one
 * large MSM sharded by base range (the only exchange step of the path, 96 / 192 bytes per GPU). */
MP_API int mp_points_sum_g1(int device, const uint8_t* points /* n x 96 */, size_t n, uint8_t out_point[MP_G1_BYTES]);
MP_API int mp_points_sum_g2(int device, const uint8_t* points /* n x 192 */, size_t n, uint8_t out_point[MP_G2_BYTES]);
/* in-place on `data` (2^log_n canonical Fr): inverse=0 fft / 1 ifft (with 1/n); coset=1 applies the g=7 coset
 * shift (coset_fft / coset_ifft of ark-poly).  Natural order in and out. */
MP_API int mp_ntt(int device, uint64_t* data, unsigned log_n, int inverse, int coset, float* out_device_ms);
/* R1CStoQAP::witness_map: h (domain_size x 4 limbs, canonical, natural order) for one assignment. */
MP_API int mp_witness_map(mp_ctx* ctx, const uint64_t* z, uint64_t* out_h);

/* ---- key-generation helper (used by the harness to build synthetic keys):
 *      out[i] = scalars[i] * G (the standard generator), uncompressed. */
MP_API int mp_fixed_base_g1(int device, const uint64_t* scalars, size_t n, uint8_t* out /* n x 96 */);
MP_API int mp_fixed_base_g2(int device, const uint64_t* scalars, size_t n, uint8_t* out /* n x 192 */);

/* Whole keys on the device.  Both write the `ProvingContext` encoding (groth16.rs:290-303) into out_pk; with out_pk == NULL they
 * only report the size in *out_len.
 *   mp_keygen: `Groth16::compile` (groth16.rs:570-586 -> ark `generate_parameters`) for a known trapdoor
 *     (tau, alpha, beta, gamma, delta; 5 x 4 limbs canonical, all non-zero) and the standard generators: Lagrange basis at tau,
 *     the QAP polynomials u_i, v_i, w_i of every variable (column sums of A, B, C), then 5n + m fixed-base multiplications.
 *     h_len = 0 selects ark's m - 1 h_query points, h_len = m the MPC form.
 *   mp_mpc_initialize: phase-2 `initialize` of the trusted setup (manta-trusted-setup/src/groth16/mpc.rs:355-431) from the phase-1
 *     powers tau^i G1 (>= 2m), tau^i G2, alpha tau^i G1, beta tau^i G1 (m each) and beta G2: h_query[i] = tau^(i+m) G - tau^i G,
 *     four group-valued inverse FFTs of size m, the sparse accumulation of :245-312, gamma = delta = 1.
 *   mp_group_ntt: the group-valued radix-2 transform alone (ark-poly `domain.fft / ifft` over curve points), in place on
 *     2^log_n uncompressed points. */
MP_API int mp_keygen(int device, const mp_r1cs_view* r1cs, const uint64_t* trapdoor /* 5 x 4 */, uint64_t h_len, uint8_t* out_pk,
                     size_t out_cap, size_t* out_len);
MP_API int mp_mpc_initialize(int device, const mp_r1cs_view* r1cs, const uint8_t* tau_powers_g1, size_t n_tau_g1,
                             const uint8_t* tau_powers_g2, const uint8_t* alpha_tau_powers_g1, const uint8_t* beta_tau_powers_g1,
                             const uint8_t* beta_g2, uint8_t* out_pk, size_t out_cap, size_t* out_len);
MP_API int mp_group_ntt(int device, int group, uint8_t* points, unsigned log_n, int inverse);

/* ---- witness-side Fr work -------------------------------------------------------------------
 * Batched Poseidon permutation over BLS12-381 Fr: `Permutation::permute` of manta-pay/src/crypto/poseidon/mod.rs:385-421,
 * 515-518 applied in place to `count` independent states of `width` elements (canonical, 4 limbs each).  round_keys:
 * (full_rounds + partial_rounds) x width in round order; mds: width x width row-major; both canonical. */
MP_API int mp_poseidon_permute(int device, int width, int full_rounds, int partial_rounds, const uint64_t* round_keys,
                               const uint64_t* mds, uint64_t* states, size_t count, float* out_device_ms);

/* ---- diagnostics (not part of the reference interface; used by the parity tests and bench) ------------------
 * Element-wise field ops on the device: field 0 = Fq (6 limbs), 1 = Fr (4 limbs); canonical in/out.
 * op: 0 add, 1 sub, 2 mul, 3 sqr(a), 4 inv(a), 5 neg(a), 6 inv(a) by the word-level binary Euclid (the batched-affine
 * MSM's inversion), 7 inv(a) by the bit-level almost-Montgomery inverse; both give the values of op 4. */
MP_API int mp_debug_field_op(int device, int field, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n);
/* Element-wise group ops on uncompressed points: group 1 = G1, 2 = G2.
 * op: 0 add (a + b), 1 double (a), 2 scalar mul (a * k[i], k = n x 4 limbs). */
MP_API int mp_debug_group_op(int device, int group, int op, const uint8_t* a, const uint8_t* b, const uint64_t* k,
                      uint8_t* out, size_t n);
/* Host-only: launch geometry of the batched-affine bucket trees for one list of n_scalars scalars with window c and `groups`
 * bucket sets (0 = one per window): the pair capacity the launcher sizes round `round` for, the number of tree levels it
 * provisions, buckets and the entry bound.  Lets the CPU tests check the capacity bound against worst-case bucket loads. */
MP_API int mp_debug_ba_geometry(int c, int groups, uint32_t n_scalars, int round, uint32_t* out_pair_cap, int* out_rounds,
                                uint32_t* out_buckets, uint32_t* out_max_entries);
/* Host-only: batched-affine round scratch of the prover for a circuit with n_vars variables and the given domain size.  A batch
 * object of `capacity` sizes the scratch for a SLAB of out[4] vectors (capacity itself up to 32, half of it above; MP_BA_SLAB_DIV) and runs a tree level of a larger live count in several launches.  out = {pair slots, thread slots} that
 * min(count, slab) vectors need, then {pair slots, thread slots} the object provides, then the slab.  The CPU tests sweep the
 * counts at the reference shapes (a partial batch re-plans its rounds and can need MORE thread slots than a full one). */
MP_API int mp_debug_prove_ba_demand(uint32_t n_vars, uint32_t domain_size, size_t capacity, size_t count, int g2, uint64_t out[5]);
/* Sustained integer-pipe rate: independent IMAD.WIDE.U32 chains on every SM; returns wide-MACs per second
 * and the measured Fq Montgomery products per second of the production multiply. */
MP_API int mp_debug_int_pipe_rate(int device, double* out_wide_mac_per_s, double* out_fq_mul_per_s);

#ifdef __cplusplus
}
#endif
#endif /* MANTAPROVER_H */
