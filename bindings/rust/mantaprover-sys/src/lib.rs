//! Raw bindings to `include/mantaprover.h` — the entry points the `ProofSystem::prove` shim needs
//! (`bindings/rust/groth16_b200.rs`).  Field order and types follow the C header exactly.
//! Written against the header by hand; NOT compiled in the build image (no Rust toolchain there).
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_float, c_int};

pub const MP_OK: c_int = 0;
pub const MP_PROOF_BYTES: usize = 192;
pub const MP_G1_BYTES: usize = 96;
pub const MP_G2_BYTES: usize = 192;

#[repr(C)]
pub struct mp_ctx {
    _private: [u8; 0],
}
#[repr(C)]
pub struct mp_batch {
    _private: [u8; 0],
}
#[repr(C)]
pub struct mp_msm_bases {
    _private: [u8; 0],
}

/// `ark_groth16::ProvingKey<Bls12_381>` as a set of borrowed pointers into its `serialize_unchecked` bytes.
#[repr(C)]
pub struct mp_pk_view {
    pub alpha_g1: *const u8,
    pub beta_g2: *const u8,
    pub gamma_g2: *const u8,
    pub delta_g2: *const u8,
    pub gamma_abc_g1: *const u8,
    pub gamma_abc_len: u64,
    pub beta_g1: *const u8,
    pub delta_g1: *const u8,
    pub a_query: *const u8,
    pub a_len: u64,
    pub b_g1_query: *const u8,
    pub b_g1_len: u64,
    pub b_g2_query: *const u8,
    pub b_g2_len: u64,
    pub h_query: *const u8,
    pub h_len: u64,
    pub l_query: *const u8,
    pub l_len: u64,
}

/// CSR form of ark-relations `ConstraintMatrices` (instance columns first), canonical coefficients.
#[repr(C)]
pub struct mp_r1cs_view {
    pub num_instance: u64,
    pub num_witness: u64,
    pub num_constraints: u64,
    pub a_row_ptr: *const u64,
    pub a_col: *const u32,
    pub a_coeff: *const u64,
    pub b_row_ptr: *const u64,
    pub b_col: *const u32,
    pub b_coeff: *const u64,
    pub c_row_ptr: *const u64,
    pub c_col: *const u32,
    pub c_coeff: *const u64,
}

extern "C" {
    pub fn mp_strerror(code: c_int) -> *const c_char;
    pub fn mp_last_error_detail() -> *const c_char;
    pub fn mp_device_count(out_count: *mut c_int) -> c_int;
    pub fn mp_pk_parse(data: *const u8, len: usize, out: *mut mp_pk_view) -> c_int;
    pub fn mp_ctx_create(pk: *const mp_pk_view, r1cs: *const mp_r1cs_view, device: c_int, out: *mut *mut mp_ctx) -> c_int;
    pub fn mp_ctx_destroy(ctx: *mut mp_ctx);
    pub fn mp_prove(ctx: *mut mp_ctx, z: *const u64, r: *const u64, s: *const u64, out_proof: *mut u8) -> c_int;
    pub fn mp_prove_from_abc(ctx: *mut mp_ctx, z: *const u64, a: *const u64, b: *const u64, c: *const u64, r: *const u64, s: *const u64,
                             out_proof: *mut u8) -> c_int;
    pub fn mp_prove_batch(ctx: *mut mp_ctx, count: usize, z: *const u64, r: *const u64, s: *const u64, out_proofs: *mut u8) -> c_int;
    pub fn mp_batch_create(ctx: *mut mp_ctx, capacity: usize, out: *mut *mut mp_batch) -> c_int;
    pub fn mp_batch_destroy(b: *mut mp_batch);
    pub fn mp_batch_submit(b: *mut mp_batch, count: usize, z: *const u64, r: *const u64, s: *const u64, out_proofs: *mut u8) -> c_int;
    pub fn mp_batch_wait(b: *mut mp_batch, out_device_ms: *mut c_float) -> c_int;
    pub fn mp_msm_g1(device: c_int, bases: *const u8, scalars: *const u64, n: usize, out_point: *mut u8, out_device_ms: *mut c_float) -> c_int;
    pub fn mp_msm_g2(device: c_int, bases: *const u8, scalars: *const u64, n: usize, out_point: *mut u8, out_device_ms: *mut c_float) -> c_int;
    pub fn mp_msm_bases_create(device: c_int, group: c_int, bases: *const u8, n: usize, out: *mut *mut mp_msm_bases) -> c_int;
    pub fn mp_msm_bases_run(h: *mut mp_msm_bases, scalars: *const u64, n: usize, out_point: *mut u8, out_device_ms: *mut c_float) -> c_int;
    pub fn mp_msm_bases_destroy(h: *mut mp_msm_bases);
    pub fn mp_keygen(device: c_int, r1cs: *const mp_r1cs_view, trapdoor: *const u64, h_len: u64, out_pk: *mut u8, out_cap: usize,
                     out_len: *mut usize) -> c_int;
    pub fn mp_mpc_initialize(device: c_int, r1cs: *const mp_r1cs_view, tau_powers_g1: *const u8, n_tau_g1: usize, tau_powers_g2: *const u8,
                             alpha_tau_powers_g1: *const u8, beta_tau_powers_g1: *const u8, beta_g2: *const u8, out_pk: *mut u8,
                             out_cap: usize, out_len: *mut usize) -> c_int;
    pub fn mp_group_ntt(device: c_int, group: c_int, points: *mut u8, log_n: u32, inverse: c_int) -> c_int;
    pub fn mp_ntt(device: c_int, data: *mut u64, log_n: u32, inverse: c_int, coset: c_int, out_device_ms: *mut c_float) -> c_int;
    pub fn mp_poseidon_permute(device: c_int, width: c_int, full_rounds: c_int, partial_rounds: c_int, round_keys: *const u64,
                               mds: *const u64, states: *mut u64, count: usize, out_device_ms: *mut c_float) -> c_int;
}
