"""Host-side randomness of the `prove` boundary (mirror of what the Rust shim does).

`Groth16::prove` (`manta-crypto/src/arkworks/groth16.rs:588-600`) hands the caller's rng, wrapped in
`SizedRng` (`manta-crypto/src/rand.rs:50-84`), to arkworks, whose `create_random_proof` draws
`r = Fr::rand(rng); s = Fr::rand(rng)` before anything else.
The signer's rng is `ChaCha20Rng` (`manta-pay/src/signer/base.rs:94`).  The C ABI takes r and s
explicitly, so this file is only needed to mirror a seeded Rust run.
"""
from __future__ import annotations

import struct

MASK32 = 0xFFFFFFFF


def _rotl(v, n):
    return ((v << n) & MASK32) | (v >> (32 - n))


def _quarter(s, a, b, c, d):
    s[a] = (s[a] + s[b]) & MASK32; s[d] = _rotl(s[d] ^ s[a], 16)
    s[c] = (s[c] + s[d]) & MASK32; s[b] = _rotl(s[b] ^ s[c], 12)
    s[a] = (s[a] + s[b]) & MASK32; s[d] = _rotl(s[d] ^ s[a], 8)
    s[c] = (s[c] + s[d]) & MASK32; s[b] = _rotl(s[b] ^ s[c], 7)


def chacha20_block(key_words, counter, stream):
    """djb ChaCha20: 64-bit block counter (words 12,13), 64-bit stream id (words 14,15)."""
    init = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + list(key_words) + [
        counter & MASK32, (counter >> 32) & MASK32, stream & MASK32, (stream >> 32) & MASK32]
    s = list(init)
    for _ in range(10):
        _quarter(s, 0, 4, 8, 12); _quarter(s, 1, 5, 9, 13); _quarter(s, 2, 6, 10, 14); _quarter(s, 3, 7, 11, 15)
        _quarter(s, 0, 5, 10, 15); _quarter(s, 1, 6, 11, 12); _quarter(s, 2, 7, 8, 13); _quarter(s, 3, 4, 9, 14)
    return [(x + y) & MASK32 for x, y in zip(s, init)]


class ChaCha20Rng:
    """rand_chacha 0.3 `ChaCha20Rng::from_seed` word stream (C.2): 64-word buffer (4 blocks),
    `next_u64` = lo | hi << 32 from two consecutive words, straddling a refill when one word is left."""

    def __init__(self, seed: bytes, stream: int = 0):
        assert len(seed) == 32
        self.key = struct.unpack("<8I", seed)
        self.stream = stream
        self.counter = 0
        self.buf = []
        self.idx = 64

    def _refill(self):
        self.buf = []
        for _ in range(4):
            self.buf.extend(chacha20_block(self.key, self.counter, self.stream))
            self.counter += 1
        self.idx = 0

    def next_u32(self):
        if self.idx >= 64:
            self._refill()
        v = self.buf[self.idx]
        self.idx += 1
        return v

    def next_u64(self):
        lo = self.next_u32()
        hi = self.next_u32()
        return lo | (hi << 32)

    def fill_bytes(self, n):
        out = bytearray()
        while len(out) < n:
            out += struct.pack("<I", self.next_u32())
        return bytes(out[:n])


def field_rand(rng, modulus: int) -> int:
    """ark-ff 0.3 `Fp::rand` (C.1): N limbs from `next_u64` (index 0 first), shave the top bits,
    reject if >= modulus; the accepted integer IS the Montgomery representation, so the sampled
    field value is limbs * R^-1 mod p.  Returns the canonical value."""
    nlimbs = (modulus.bit_length() + 63) // 64
    shave = 64 * nlimbs - modulus.bit_length()
    while True:
        limbs = [rng.next_u64() for _ in range(nlimbs)]
        limbs[-1] &= (1 << (64 - shave)) - 1
        v = sum(l << (64 * i) for i, l in enumerate(limbs))
        if v < modulus:
            return v * pow(1 << (64 * nlimbs), -1, modulus) % modulus
