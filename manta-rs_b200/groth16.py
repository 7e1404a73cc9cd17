"""Host-side mirror of the reference's Groth16 plugin interface, backed by the CUDA library.

Mirrors `manta-crypto/src/arkworks/groth16.rs` (names and argument meaning):
  Error (:50-60), Proof (:62-81, bytes :184-195, codec::Encode :159-170), ProvingContext (:208-303),
  Groth16::prove (:588-600) — and the compiler `R1CS<F>` of `constraint/mod.rs:91-217` as far as `prove`
  consumes it (the finalized matrices and the full assignment).
The reference toolchain (Rust) is absent from this image, so this Python layer plays the role of the Rust shim
described in INTEGRATION.md: draw r, s from the caller's rng exactly like ark-groth16's `create_random_proof`,
then make ONE call into the C ABI.  All arithmetic of the path runs in libmantaprover.so on the GPU.
"""
from __future__ import annotations

import ctypes
import hashlib

from . import _native as nat
from .rng import field_rand
from .workload import FR_BLS12_381


class Error(Exception):
    """The opaque unit error of groth16.rs:50-60 (every failure collapses to it)."""


class Proof:
    """Compressed a | b | c, 192 bytes on BLS12-381 (`proof_as_bytes`, groth16.rs:184-195)."""

    SIZE = nat.PROOF_BYTES

    def __init__(self, data: bytes):
        if len(data) != self.SIZE:
            raise Error()
        self.data = bytes(data)

    def to_bytes(self) -> bytes:
        return self.data

    def encode(self) -> bytes:
        """`codec::Encode` for Proof (groth16.rs:159-170): the bytes as a Vec<u8>, i.e. u64-LE length prefix
        (`manta-util/src/codec.rs:672-686`)."""
        return len(self.data).to_bytes(8, "little") + self.data

    @classmethod
    def try_from_bytes(cls, data: bytes) -> "Proof":
        return cls(data)

    def __eq__(self, other):
        return isinstance(other, Proof) and self.data == other.data

    def __hash__(self):
        return hash(self.data)


def _csr(rows):
    row_ptr, cols, coeffs = [0], [], []
    for row in rows:
        for coeff, col in row:
            cols.append(col)
            coeffs.append(coeff)
        row_ptr.append(len(cols))
    rp = (ctypes.c_uint64 * len(row_ptr))(*row_ptr)
    cl = (ctypes.c_uint32 * max(len(cols), 1))(*cols)
    cf = ctypes.create_string_buffer(nat.pack_scalars(coeffs), max(len(coeffs), 1) * 32)
    return rp, cl, cf


class ConstraintMatrices:
    """The per-circuit constant part of a finalized constraint system: ark `to_matrices()` output."""

    def __init__(self, num_instance: int, num_witness: int, a, b, c):
        assert len(a) == len(b) == len(c)
        self.p, self.w, self.K = num_instance, num_witness, len(a)
        self.a, self.b, self.c = a, b, c
        self._csr = None
        self._digest = None

    @property
    def n(self):
        return self.p + self.w

    def view(self):
        if self._csr is None:
            self._csr = [_csr(m) for m in (self.a, self.b, self.c)]
        v = nat.R1csView()
        v.num_instance, v.num_witness, v.num_constraints = self.p, self.w, self.K
        for name, (rp, cl, cf) in zip("abc", self._csr):
            setattr(v, f"{name}_row_ptr", ctypes.cast(rp, ctypes.c_void_p))
            setattr(v, f"{name}_col", ctypes.cast(cl, ctypes.c_void_p))
            setattr(v, f"{name}_coeff", ctypes.cast(cf, ctypes.c_void_p))
        return v

    def digest(self) -> bytes:
        if self._digest is None:
            h = hashlib.blake2b(digest_size=16)
            h.update(f"{self.p},{self.w},{self.K}".encode())
            for m in (self.a, self.b, self.c):
                for row in m:
                    h.update(repr(row).encode())
            self._digest = h.digest()
        return self._digest


class R1CS:
    """`R1CS<F>` as handed to `prove`: already synthesized (constraint/mod.rs:199-217 moves the pre-built system
    into arkworks' container) — matrices plus the full assignment z = [1, instance.., witness..]."""

    def __init__(self, matrices: ConstraintMatrices, assignment):
        if len(assignment) != matrices.n or assignment[0] != 1:
            raise Error()
        self.matrices = matrices
        self.assignment = assignment

    @classmethod
    def from_workload(cls, cs, z):
        """Adapter for `workload.R1CS` synthetic systems."""
        if not hasattr(cs, "_matrices"):
            cs._matrices = ConstraintMatrices(cs.p, cs.w, cs.a, cs.b, cs.c)
        return cls(cs._matrices, z)


class ProvingContext:
    """`ProvingContext<E>` (groth16.rs:208-303): owns the proving key; `decode`/`encode` use the reference's
    on-disk format (`serialize_unchecked`: uncompressed points, u64-LE vector lengths).  Device residency is
    created lazily per (device, circuit) and shared by clones, like the `Arc` SURVEY.md §3.5 asks for."""

    def __init__(self, pk_bytes: bytes):
        self.pk_bytes = bytes(pk_bytes)
        self._buf = ctypes.create_string_buffer(self.pk_bytes, len(self.pk_bytes))
        self._view = nat.PkView()
        try:
            nat.check(nat.lib().mp_pk_parse(self._buf, len(self.pk_bytes), ctypes.byref(self._view)))
        except nat.NativeError as e:
            raise Error() from e
        self._native = {}

    @classmethod
    def decode(cls, data: bytes) -> "ProvingContext":
        return cls(data)

    def encode(self) -> bytes:
        return self.pk_bytes

    def clone(self) -> "ProvingContext":
        return self  # immutable; device state is shared

    def __eq__(self, other):
        return isinstance(other, ProvingContext) and self.pk_bytes == other.pk_bytes

    def __hash__(self):
        return hash(self.pk_bytes)

    def native(self, matrices: ConstraintMatrices, device: int = 0):
        key = (device, matrices.digest())
        h = self._native.get(key)
        if h is None:
            h = ctypes.c_void_p()
            v = matrices.view()
            nat.check(nat.lib().mp_ctx_create(ctypes.byref(self._view), ctypes.byref(v), device, ctypes.byref(h)))
            self._native[key] = h
        return h

    def close(self):
        for h in self._native.values():
            nat.lib().mp_ctx_destroy(h)
        self._native = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Groth16:
    """`Groth16<Bls12_381>` as a `ProofSystem` (groth16.rs:548-610) — prove only; compile/verify stay on the
    reference's CPU path."""

    MODULUS = FR_BLS12_381
    device = 0

    @staticmethod
    def draw_randomness(rng):
        """`create_random_proof`: r = Fr::rand(rng); s = Fr::rand(rng), in that order, before anything else."""
        r = field_rand(rng, Groth16.MODULUS)
        s = field_rand(rng, Groth16.MODULUS)
        return r, s

    @classmethod
    def prove(cls, context: ProvingContext, compiler: R1CS, rng) -> Proof:
        r, s = cls.draw_randomness(rng)
        return cls.prove_with_randomness(context, compiler, r, s)

    @classmethod
    def prove_with_randomness(cls, context: ProvingContext, compiler: R1CS, r: int, s: int) -> Proof:
        try:
            h = context.native(compiler.matrices, cls.device)
            out = ctypes.create_string_buffer(nat.PROOF_BYTES)
            nat.check(nat.lib().mp_prove(h, nat.pack_scalars(compiler.assignment), nat.pack_scalars([r]),
                                         nat.pack_scalars([s]), out))
            return Proof(out.raw)
        except nat.NativeError as e:
            raise Error() from e

    @classmethod
    def prove_many(cls, context: ProvingContext, compilers, rng):
        """Batch extension: same results as looping `prove` with the same rng."""
        rs = [cls.draw_randomness(rng) for _ in compilers]
        return cls.prove_many_with_randomness(context, compilers, [x[0] for x in rs], [x[1] for x in rs])

    @classmethod
    def prove_many_with_randomness(cls, context, compilers, rs, ss):
        if not compilers:
            return []
        mats = compilers[0].matrices
        if any(c.matrices is not mats for c in compilers):
            raise Error()
        try:
            h = context.native(mats, cls.device)
            z = b"".join(nat.pack_scalars(c.assignment) for c in compilers)
            out = ctypes.create_string_buffer(nat.PROOF_BYTES * len(compilers))
            nat.check(nat.lib().mp_prove_batch(h, len(compilers), z, nat.pack_scalars(rs), nat.pack_scalars(ss), out))
            return [Proof(out.raw[i * nat.PROOF_BYTES:(i + 1) * nat.PROOF_BYTES]) for i in range(len(compilers))]
        except nat.NativeError as e:
            raise Error() from e


class ProverPool:
    """Batch proving over several GPUs of one process with the fault path of SURVEY.md §5: the batch is cut into chunks, every
    device (one worker thread each, the C ABI call releases the GIL) pulls chunks from one queue, and when a device fails
    (any non-zero return code of the library: CUDA error, out of memory, lost device) it is retired and its chunk is
    re-queued for the surviving devices.  Proofs are independent units, so the result does not depend on
    which device produced which proof.  Raises `Error` only when every device has failed."""

    def __init__(self, context: ProvingContext, devices, chunk: int = 128, prove_chunk=None):
        self.context = context
        self.devices = list(devices)
        self.chunk = max(1, chunk)
        self.failed = {}            # device -> the error that retired it
        self._prove_chunk = prove_chunk or self._native_chunk

    def _native_chunk(self, device, compilers, rs, ss):
        mats = compilers[0].matrices
        h = self.context.native(mats, device)
        z = b"".join(nat.pack_scalars(c.assignment) for c in compilers)
        out = ctypes.create_string_buffer(nat.PROOF_BYTES * len(compilers))
        nat.check(nat.lib().mp_prove_batch(h, len(compilers), z, nat.pack_scalars(rs), nat.pack_scalars(ss), out))
        return [out.raw[i * nat.PROOF_BYTES:(i + 1) * nat.PROOF_BYTES] for i in range(len(compilers))]

    def prove_many_with_randomness(self, compilers, rs, ss):
        import queue
        import threading
        n = len(compilers)
        if n == 0:
            return []
        if any(c.matrices is not compilers[0].matrices for c in compilers):
            raise Error()
        work = queue.Queue()
        for lo in range(0, n, self.chunk):
            work.put(list(range(lo, min(n, lo + self.chunk))))
        results = [None] * n
        lock = threading.Lock()
        pending = [work.qsize()]

        def worker(device):
            while True:
                with lock:
                    if pending[0] == 0 or device in self.failed:
                        return
                try:
                    idx = work.get(timeout=0.05)
                except queue.Empty:
                    continue
                try:
                    proofs = self._prove_chunk(device, [compilers[i] for i in idx], [rs[i] for i in idx], [ss[i] for i in idx])
                    if len(proofs) != len(idx) or any(len(p) != nat.PROOF_BYTES for p in proofs):
                        raise nat.NativeError(2, "malformed result", f"device {device}")
                except nat.NativeError as e:
                    with lock:
                        self.failed[device] = e
                    work.put(idx)           # re-queue for the surviving devices
                    return
                with lock:
                    for i, p in zip(idx, proofs):
                        results[i] = Proof(p)
                    pending[0] -= 1

        live = [d for d in self.devices if d not in self.failed]
        threads = [threading.Thread(target=worker, args=(d,), daemon=True) for d in live]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if any(r is None for r in results):
            raise Error()           # every device failed
        return results
