//! `ProofSystem::prove` for `Groth16<Bls12_381>` over libmantaprover.so — the shim of INTEGRATION.md §2 as a source file.
//! Drop into manta-crypto/src/arkworks/ behind `feature = "b200"`; depends on `bindings/rust/mantaprover-sys`.
//! NOT compiled in the build image (no Rust toolchain there); the C++ and Python host mirrors of this repository implement
//! the same steps and are tested bit-exact on the GPU (tests/test_cpp_host.py, tests/test_gpu_parity.py).
// manta-crypto/src/arkworks/groth16_b200.rs  (feature = "b200")
use ark_ff::{PrimeField, UniformRand, BigInteger};
use ark_relations::r1cs::{ConstraintSynthesizer, ConstraintSystem, OptimizationGoal};
use std::sync::Arc;

#[repr(C)] pub struct mp_ctx { _p: [u8; 0] }
#[repr(C)] pub struct mp_pk_view { /* field order of include/mantaprover.h */ }
#[repr(C)] pub struct mp_r1cs_view { /* ... */ }
extern "C" {
    fn mp_pk_parse(data: *const u8, len: usize, out: *mut mp_pk_view) -> i32;
    fn mp_ctx_create(pk: *const mp_pk_view, r1cs: *const mp_r1cs_view, device: i32, out: *mut *mut mp_ctx) -> i32;
    fn mp_ctx_destroy(ctx: *mut mp_ctx);
    fn mp_prove(ctx: *mut mp_ctx, z: *const u64, r: *const u64, s: *const u64, out: *mut u8) -> i32;
    fn mp_prove_batch(ctx: *mut mp_ctx, count: usize, z: *const u64, r: *const u64, s: *const u64, out: *mut u8) -> i32;
}

pub struct DeviceContext(*mut mp_ctx);
unsafe impl Send for DeviceContext {}
unsafe impl Sync for DeviceContext {}
impl Drop for DeviceContext { fn drop(&mut self) { unsafe { mp_ctx_destroy(self.0) } } }

/// `ProvingContext<E>` keeps its `ark_groth16::ProvingKey<E>` (so `Clone/Eq/Hash/Encode/Decode` are unchanged) plus a
/// lazily created, shared device context keyed by the circuit's matrix digest.
pub struct ProvingContext<E: PairingEngine> {
    pub proving_key: ProvingKey<E>,
    device: OnceCell<Arc<DeviceContext>>,
}

impl ProofSystem for Groth16<Bls12_381> {
    // ... associated types, compile, verify exactly as in groth16.rs ...
    fn prove<R>(context: &Self::ProvingContext, compiler: Self::Compiler, rng: &mut R) -> Result<Self::Proof, Self::Error>
    where R: CryptoRng + RngCore + ?Sized,
    {
        // (1) ark_groth16::create_random_proof: r then s, before anything else
        let mut rng = SizedRng(rng);
        let r = Fr::rand(&mut rng);
        let s = Fr::rand(&mut rng);
        // (2) what create_proof does before its arithmetic: move the pre-built system in, inline LCs
        let cs = ConstraintSystem::new_ref();
        cs.set_optimization_goal(OptimizationGoal::Constraints);
        compiler.generate_constraints(cs.clone()).map_err(|_| Error)?;       // constraint/mod.rs:199-217
        cs.finalize();
        let cs = cs.borrow().ok_or(Error)?;
        // (3) one-time: matrices + key to the device
        let dev = context.device.get_or_try_init(|| {
            let m = cs.to_matrices().ok_or(Error)?;                            // CSR flattening omitted
            let mut pk_bytes = Vec::new();
            context.proving_key.serialize_unchecked(&mut pk_bytes).map_err(|_| Error)?;
            let mut view = MaybeUninit::uninit();
            check(unsafe { mp_pk_parse(pk_bytes.as_ptr(), pk_bytes.len(), view.as_mut_ptr()) })?;
            let mut ctx = core::ptr::null_mut();
            check(unsafe { mp_ctx_create(view.as_ptr(), &csr_view(&m), 0, &mut ctx) })?;
            Ok(Arc::new(DeviceContext(ctx)))
        })?;
        // (4) every proof: full assignment in canonical limbs, one FFI call
        let z: Vec<u64> = cs.instance_assignment.iter().chain(cs.witness_assignment.iter())
            .flat_map(|x| x.into_repr().0).collect();
        let mut out = [0u8; 192];
        check(unsafe { mp_prove(dev.0, z.as_ptr(), r.into_repr().0.as_ptr(), s.into_repr().0.as_ptr(), out.as_mut_ptr()) })?;
        // (5) `Proof<E>` from its canonical bytes (groth16.rs:63-72, TryFrom<Vec<u8>>)
        Proof::try_from(out.to_vec()).map_err(|_| Error)
    }
}
fn check(rc: i32) -> Result<(), Error> { if rc == 0 { Ok(()) } else { Err(Error) } }
