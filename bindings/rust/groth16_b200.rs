//! `ProofSystem::prove` of `Groth16<E>` over libmantaprover.so — drop-in module for
//! `manta-crypto/src/arkworks/groth16_b200.rs`, enabled by `feature = "b200"` (see `patches/`).
//!
//! The reference implements `impl<E> ProofSystem for Groth16<E>` generically (`groth16.rs:548-610`), so a second
//! `impl ProofSystem for Groth16<Bls12_381>` would overlap it.  This module therefore stays generic itself and is called FROM
//! the generic `prove` (`patches/groth16_rs.patch`): `supports::<E>()` is a `TypeId` comparison, everything else only uses
//! the byte-level traits every `PairingEngine` offers (`CanonicalSerialize`, `into_repr`).  `ProvingContext<E>` is not
//! touched (its `CanonicalSerialize` / serde derives stay as they are): device contexts live in a process-wide registry
//! keyed by a fingerprint of the proving key.
//!
//! NOT compiled in the build image (no Rust toolchain there).  The C++ and Python host mirrors of this repository implement
//! the same steps and are tested bit-exact on the GPU (tests/test_cpp_host.py, tests/test_gpu_parity.py).

use crate::{
    arkworks::{
        bls12_381::Bls12_381,
        constraint::R1CS,
        ec::PairingEngine,
        ff::{PrimeField, UniformRand},
        groth16::Error,
        relations::r1cs::{
            ConstraintMatrices, ConstraintSynthesizer, ConstraintSystem, ConstraintSystemRef,
            OptimizationGoal,
        },
        serialize::{CanonicalDeserialize, CanonicalSerialize},
    },
    rand::{CryptoRng, RngCore, SizedRng},
};
use ark_groth16::{Proof, ProvingKey};
use core::any::TypeId;
use mantaprover_sys::{
    mp_ctx, mp_ctx_create, mp_ctx_destroy, mp_pk_parse, mp_pk_view, mp_prove, mp_prove_batch,
    mp_r1cs_view, MP_OK, MP_PROOF_BYTES,
};
use std::{
    collections::{hash_map::DefaultHasher, HashMap},
    hash::{Hash, Hasher},
    mem::MaybeUninit,
    sync::{Arc, Mutex, OnceLock},
};

/// Returns `true` when the device backend handles the pairing engine `E` (BLS12-381 only).
#[inline]
pub fn supports<E>() -> bool
where
    E: PairingEngine,
{
    TypeId::of::<E>() == TypeId::of::<Bls12_381>()
}

/// Owned device context (proving key tables + circuit matrices resident on one GPU).
struct DeviceContext(*mut mp_ctx);

// SAFETY: libmantaprover serializes calls on one context internally (include/mantaprover.h, "Conventions").
unsafe impl Send for DeviceContext {}
unsafe impl Sync for DeviceContext {}

impl Drop for DeviceContext {
    #[inline]
    fn drop(&mut self) {
        // SAFETY: the pointer came from `mp_ctx_create` and is dropped exactly once.
        unsafe { mp_ctx_destroy(self.0) }
    }
}

/// Process-wide registry of device contexts, keyed by [`fingerprint`].
fn registry() -> &'static Mutex<HashMap<u64, Arc<DeviceContext>>> {
    static REGISTRY: OnceLock<Mutex<HashMap<u64, Arc<DeviceContext>>>> = OnceLock::new();
    REGISTRY.get_or_init(Default::default)
}

/// Cheap fingerprint of a proving key: the six single points (delta is unique per key), the query lengths and the
/// first and last `a_query` points.  Hashing the whole key (`impl Hash for ProvingContext`) would cost ~23 MB per proof.
fn fingerprint<E>(pk: &ProvingKey<E>) -> u64
where
    E: PairingEngine,
{
    let mut hasher = DefaultHasher::new();
    pk.vk.alpha_g1.hash(&mut hasher);
    pk.vk.beta_g2.hash(&mut hasher);
    pk.vk.gamma_g2.hash(&mut hasher);
    pk.vk.delta_g2.hash(&mut hasher);
    pk.beta_g1.hash(&mut hasher);
    pk.delta_g1.hash(&mut hasher);
    (pk.a_query.len(), pk.h_query.len(), pk.l_query.len()).hash(&mut hasher);
    pk.a_query.first().hash(&mut hasher);
    pk.a_query.last().hash(&mut hasher);
    hasher.finish()
}

/// CSR copy of one constraint matrix with canonical little-endian coefficients (`mp_r1cs_view` layout).
struct Csr {
    row_ptr: Vec<u64>,
    col: Vec<u32>,
    coeff: Vec<u64>,
}

fn csr<F>(rows: &[Vec<(F, usize)>]) -> Csr
where
    F: PrimeField,
{
    let mut out = Csr {
        row_ptr: Vec::with_capacity(rows.len() + 1),
        col: Vec::new(),
        coeff: Vec::new(),
    };
    out.row_ptr.push(0);
    for row in rows {
        for (coeff, index) in row {
            out.col.push(*index as u32);
            out.coeff.extend_from_slice(coeff.into_repr().as_ref());
        }
        out.row_ptr.push(out.col.len() as u64);
    }
    out
}

/// Builds the `mp_r1cs_view` over three [`Csr`] matrices; the view borrows them.
fn csr_view<F>(matrices: &ConstraintMatrices<F>, a: &Csr, b: &Csr, c: &Csr) -> mp_r1cs_view
where
    F: PrimeField,
{
    mp_r1cs_view {
        num_instance: matrices.num_instance_variables as u64,
        num_witness: matrices.num_witness_variables as u64,
        num_constraints: matrices.num_constraints as u64,
        a_row_ptr: a.row_ptr.as_ptr(),
        a_col: a.col.as_ptr(),
        a_coeff: a.coeff.as_ptr(),
        b_row_ptr: b.row_ptr.as_ptr(),
        b_col: b.col.as_ptr(),
        b_coeff: b.coeff.as_ptr(),
        c_row_ptr: c.row_ptr.as_ptr(),
        c_col: c.col.as_ptr(),
        c_coeff: c.coeff.as_ptr(),
    }
}

#[inline]
fn check(code: i32) -> Result<(), Error> {
    if code == MP_OK {
        Ok(())
    } else {
        Err(Error)
    }
}

/// Returns the device context of `pk`, creating it (key upload, Montgomery conversion, window tables) on first use.
fn device_context<E>(
    pk: &ProvingKey<E>,
    matrices: impl FnOnce() -> Option<ConstraintMatrices<E::Fr>>,
) -> Result<Arc<DeviceContext>, Error>
where
    E: PairingEngine,
{
    let key = fingerprint(pk);
    let mut registry = registry().lock().map_err(|_| Error)?;
    if let Some(context) = registry.get(&key) {
        return Ok(context.clone());
    }
    let matrices = matrices().ok_or(Error)?;
    let (a, b, c) = (csr(&matrices.a), csr(&matrices.b), csr(&matrices.c));
    let r1cs = csr_view(&matrices, &a, &b, &c);
    let mut pk_bytes = Vec::new();
    pk.serialize_unchecked(&mut pk_bytes).map_err(|_| Error)?;
    let mut view = MaybeUninit::<mp_pk_view>::uninit();
    // SAFETY: `pk_bytes` outlives both calls; `mp_pk_parse` fills `view` completely on success.
    let context = unsafe {
        check(mp_pk_parse(pk_bytes.as_ptr(), pk_bytes.len(), view.as_mut_ptr()))?;
        let mut raw = core::ptr::null_mut();
        check(mp_ctx_create(view.as_ptr(), &r1cs, 0, &mut raw))?;
        Arc::new(DeviceContext(raw))
    };
    registry.insert(key, context.clone());
    Ok(context)
}

/// What `ark_groth16::create_proof` does before its arithmetic: move the pre-built system in
/// (`constraint/mod.rs:199-217`), inline the linear combinations, and read off the full assignment.
fn synthesize<F>(compiler: R1CS<F>) -> Result<ConstraintSystemRef<F>, Error>
where
    F: PrimeField,
{
    let cs = ConstraintSystem::new_ref();
    cs.set_optimization_goal(OptimizationGoal::Constraints);
    compiler.generate_constraints(cs.clone()).map_err(|_| Error)?;
    cs.finalize();
    Ok(cs)
}

fn assignment_limbs<F>(cs: &ConstraintSystem<F>, out: &mut Vec<u64>)
where
    F: PrimeField,
{
    for value in cs.instance_assignment.iter().chain(cs.witness_assignment.iter()) {
        out.extend_from_slice(value.into_repr().as_ref());
    }
}

/// Device form of `ArkGroth16::prove(&context.proving_key, compiler, &mut SizedRng(rng))` (`groth16.rs:597`).
pub fn prove<E, R>(
    pk: &ProvingKey<E>,
    compiler: R1CS<E::Fr>,
    rng: &mut R,
) -> Result<Proof<E>, Error>
where
    E: PairingEngine,
    R: CryptoRng + RngCore + ?Sized,
{
    // `create_random_proof`: r then s, before anything else.
    let mut rng = SizedRng(rng);
    let r = E::Fr::rand(&mut rng);
    let s = E::Fr::rand(&mut rng);
    let cs = synthesize(compiler)?;
    let context = device_context(pk, || cs.to_matrices())?;
    let cs = cs.borrow().ok_or(Error)?;
    let mut z = Vec::with_capacity(4 * (cs.num_instance_variables + cs.num_witness_variables));
    assignment_limbs(&cs, &mut z);
    let mut out = [0u8; MP_PROOF_BYTES];
    // SAFETY: every buffer outlives the call; sizes follow include/mantaprover.h.
    check(unsafe {
        mp_prove(
            context.0,
            z.as_ptr(),
            r.into_repr().as_ref().as_ptr(),
            s.into_repr().as_ref().as_ptr(),
            out.as_mut_ptr(),
        )
    })?;
    Proof::deserialize(&out[..]).map_err(|_| Error)
}

/// Batch form behind `ProofSystem::prove_many` (`patches/constraint_rs.patch`): the same proofs as calling [`prove`] once
/// per compiler with the same `rng`, in ONE device batch.
pub fn prove_many<E, R>(
    pk: &ProvingKey<E>,
    compilers: Vec<R1CS<E::Fr>>,
    rng: &mut R,
) -> Result<Vec<Proof<E>>, Error>
where
    E: PairingEngine,
    R: CryptoRng + RngCore + ?Sized,
{
    let count = compilers.len();
    let (mut z, mut rs, mut ss) = (Vec::new(), Vec::new(), Vec::new());
    let mut context = None;
    for compiler in compilers {
        // Same rng consumption as `count` sequential `prove` calls: two draws per proof, nothing in between.
        let mut sized = SizedRng(&mut *rng);
        rs.extend_from_slice(E::Fr::rand(&mut sized).into_repr().as_ref());
        ss.extend_from_slice(E::Fr::rand(&mut sized).into_repr().as_ref());
        let cs = synthesize(compiler)?;
        if context.is_none() {
            context = Some(device_context(pk, || cs.to_matrices())?);
        }
        assignment_limbs(&*cs.borrow().ok_or(Error)?, &mut z);
    }
    let context = match context {
        Some(context) => context,
        None => return Ok(Vec::new()),
    };
    let mut out = vec![0u8; count * MP_PROOF_BYTES];
    // SAFETY: as in `prove`.
    check(unsafe {
        mp_prove_batch(context.0, count, z.as_ptr(), rs.as_ptr(), ss.as_ptr(), out.as_mut_ptr())
    })?;
    out.chunks_exact(MP_PROOF_BYTES)
        .map(|bytes| Proof::deserialize(bytes).map_err(|_| Error))
        .collect()
}
