"""Proving keys built on the device.

The reference generates its benchmark keys with `Groth16::compile` (groth16.rs:570-586 -> ark `circuit_specific_setup`)
from a seeded rng (`manta-pay/src/parameters.rs:56-106`) and its production keys with the phase-2 `initialize` of the
trusted setup (`manta-trusted-setup/src/groth16/mpc.rs:355-431`).  Both run in libmantaprover.so here: the QAP evaluation at
tau, the 5n + m fixed-base multiplications, the group-valued inverse FFTs and the sparse accumulation are CUDA kernels
(csrc/keygen.cu); this module only marshals the constraint matrices and the byte strings.  Output: the reference's
`ProvingContext` byte format (groth16.rs:290-303).
"""
from __future__ import annotations

import ctypes

from . import _native as nat
from .groth16 import ConstraintMatrices


def _matrices(cs) -> ConstraintMatrices:
    if isinstance(cs, ConstraintMatrices):
        return cs
    if not hasattr(cs, "_matrices"):
        cs._matrices = ConstraintMatrices(cs.p, cs.w, cs.a, cs.b, cs.c)
    return cs._matrices


def generate(cs, trapdoor, device: int = 0, h_len=None) -> bytes:
    """`ProvingContext` bytes for the trapdoor (tau, alpha, beta, gamma, delta) and the standard generators."""
    view = _matrices(cs).view()
    trap = nat.pack_scalars(list(trapdoor))
    size = ctypes.c_size_t()
    lib = nat.lib()
    nat.check(lib.mp_keygen(device, ctypes.byref(view), trap, h_len or 0, None, 0, ctypes.byref(size)))
    out = ctypes.create_string_buffer(size.value)
    nat.check(lib.mp_keygen(device, ctypes.byref(view), trap, h_len or 0, out, size.value, ctypes.byref(size)))
    return out.raw


def mpc_initialize(cs, tau_powers_g1: bytes, tau_powers_g2: bytes, alpha_tau_powers_g1: bytes, beta_tau_powers_g1: bytes,
                   beta_g2: bytes, device: int = 0) -> bytes:
    """Phase-2 `initialize` (mpc.rs:355-431): `ProvingContext` bytes of the initial MPC state (gamma = delta = 1,
    m h_query points) from the phase-1 accumulator powers (ark uncompressed points)."""
    view = _matrices(cs).view()
    size = ctypes.c_size_t()
    lib = nat.lib()
    nat.check(lib.mp_mpc_initialize(device, ctypes.byref(view), None, 0, None, None, None, None, None, 0, ctypes.byref(size)))
    out = ctypes.create_string_buffer(size.value)
    nat.check(lib.mp_mpc_initialize(device, ctypes.byref(view), tau_powers_g1, len(tau_powers_g1) // nat.G1_BYTES, tau_powers_g2,
                                    alpha_tau_powers_g1, beta_tau_powers_g1, beta_g2, out, size.value, ctypes.byref(size)))
    return out.raw


def group_ntt(group: int, points: bytes, inverse: bool, device: int = 0) -> bytes:
    """ark-poly `domain.fft` / `domain.ifft` over 2^k curve points (ark uncompressed in and out)."""
    pb = nat.G1_BYTES if group == 1 else nat.G2_BYTES
    n = len(points) // pb
    assert n and n & (n - 1) == 0 and n * pb == len(points)
    buf = ctypes.create_string_buffer(points, len(points))
    nat.check(nat.lib().mp_group_ntt(device, group, buf, n.bit_length() - 1, 1 if inverse else 0))
    return buf.raw
