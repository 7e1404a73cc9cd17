"""Host mirror of manta-pay's Poseidon permutation / hasher for batches of independent inputs.

Mirrors `manta-pay/src/crypto/poseidon/mod.rs` (`Permutation { additive_round_keys, mds_matrix }`, `permute`, :315-421,
:515-518) and `hash.rs:113-127` (`Hasher::hash_untruncated`: state = domain_tag ‖ inputs, permute, output element 0), over
BLS12-381 Fr, on the CUDA path `mp_poseidon_permute` — no CPU fallback.
"""
from __future__ import annotations

import ctypes

from . import _native as nat


class Permutation:
    def __init__(self, width, full_rounds, partial_rounds, additive_round_keys, mds_matrix, device=0):
        if len(additive_round_keys) != (full_rounds + partial_rounds) * width or len(mds_matrix) != width * width:
            raise ValueError("round keys / MDS matrix do not match the width and round numbers")
        self.width, self.full_rounds, self.partial_rounds, self.device = width, full_rounds, partial_rounds, device
        self._rk = nat.pack_scalars(list(additive_round_keys))
        self._mds = nat.pack_scalars(list(mds_matrix))
        self.last_device_ms = 0.0

    def permute_many(self, states):
        """states: list of `width`-element lists of canonical integers; returns the permuted states."""
        count = len(states)
        if any(len(s) != self.width for s in states):
            raise ValueError("state width mismatch")
        buf = ctypes.create_string_buffer(nat.pack_scalars([x for s in states for x in s]), max(1, count * self.width * 32))
        ms = ctypes.c_float()
        nat.check(nat.lib().mp_poseidon_permute(self.device, self.width, self.full_rounds, self.partial_rounds, self._rk, self._mds,
                                                buf, count, ctypes.byref(ms)))
        self.last_device_ms = ms.value
        flat = nat.unpack_scalars(buf.raw[:count * self.width * 32])
        return [flat[i * self.width:(i + 1) * self.width] for i in range(count)]

    def permute(self, state):
        return self.permute_many([state])[0]


class Hasher:
    """`Hasher<S, T, ARITY>`: hash(inputs) = permute(domain_tag ‖ inputs)[0] (hash.rs:113-127, :150-158)."""

    def __init__(self, permutation: Permutation, domain_tag: int):
        self.permutation, self.domain_tag = permutation, domain_tag

    @property
    def arity(self):
        return self.permutation.width - 1

    def hash_many(self, inputs):
        return [s[0] for s in self.permutation.permute_many([[self.domain_tag] + list(x) for x in inputs])]

    def hash(self, inputs):
        return self.hash_many([inputs])[0]
