// C++ host side of the drop-in boundary: the reference's Groth16 plugin interface over the C ABI of mantaprover.h.
//
// The reference's host language is Rust and no Rust toolchain exists in the build image, so this header plays the role
// of the Rust shim of INTEGRATION.md for compiled-code callers.  It mirrors, name for name and with the same argument
// meaning and error behaviour, manta-crypto/src/arkworks/groth16.rs:
//   Error (:50-60, the opaque unit error every failure collapses to)      -> manta::groth16::Error
//   Proof<E> (:62-81; bytes :184-195; codec::Encode :159-170)              -> manta::groth16::Proof
//   ProvingContext<E> (:208-303; Decode/Encode :268-303)                   -> manta::groth16::ProvingContext
//   R1CS<F> as consumed by prove (constraint/mod.rs:91-217)                 -> manta::groth16::R1CS + ConstraintMatrices
//   Groth16::prove(context, compiler, rng) (:588-600)                      -> manta::groth16::Groth16::prove
// and the randomness rule of ark-groth16 0.3 `create_random_proof` behind it: r = Fr::rand(rng), s = Fr::rand(rng), drawn
// first, with ark-ff 0.3 `Fp::rand` (four `next_u64` limbs, top bit shaved, rejection, the accepted integer being the
// Montgomery representation) over the signer's `ChaCha20Rng` (manta-pay/src/signer/base.rs:94; rand_chacha 0.3 word stream).
// Header-only, C++17; link with -lmantaprover.  All arithmetic of the path runs in the CUDA library; there is no CPU
// fallback: without an sm_100 device every prove returns Error.
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <utility>
#include <variant>
#include <vector>

#include "mantaprover.h"

namespace manta::groth16 {

// ---- Error / Result -----------------------------------------------------------------------------------------------
struct Error {};  // groth16.rs:50-60: deliberately carries nothing

template <class T>
class Result {
   public:
    Result(T v) : v_(std::move(v)) {}
    Result(Error e) : v_(e) {}
    bool is_ok() const { return v_.index() == 0; }
    bool is_err() const { return !is_ok(); }
    const T& unwrap() const { return std::get<0>(v_); }
    T& unwrap() { return std::get<0>(v_); }

   private:
    std::variant<T, Error> v_;
};

// ---- scalars --------------------------------------------------------------------------------------------------------
using Fr = std::array<uint64_t, 4>;  // canonical value, little-endian limbs (`into_repr()`)

namespace detail {
constexpr uint64_t FR_MOD[4] = {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull};
constexpr uint64_t FR_INV = 0xfffffffeffffffffull;  // -r^-1 mod 2^64

inline bool geq_mod(const uint64_t v[4]) {
    for (int i = 3; i >= 0; i--) {
        if (v[i] > FR_MOD[i]) return true;
        if (v[i] < FR_MOD[i]) return false;
    }
    return true;
}
// v * 2^-256 mod r: the canonical value of a Montgomery representation
inline Fr from_montgomery(const Fr& v) {
    uint64_t t[9] = {v[0], v[1], v[2], v[3], 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        const uint64_t m = t[i] * FR_INV;
        unsigned __int128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (unsigned __int128)m * FR_MOD[j] + t[i + j];
            t[i + j] = (uint64_t)c;
            c >>= 64;
        }
        for (int j = i + 4; j < 9 && c; j++) {
            c += t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
    }
    uint64_t r[4] = {t[4], t[5], t[6], t[7]};
    if (geq_mod(r)) {
        unsigned __int128 b = 0;
        for (int j = 0; j < 4; j++) {
            unsigned __int128 d = (unsigned __int128)r[j] - FR_MOD[j] - (uint64_t)b;
            r[j] = (uint64_t)d;
            b = (d >> 64) & 1;
        }
    }
    return {r[0], r[1], r[2], r[3]};
}
}  // namespace detail

// ---- rng --------------------------------------------------------------------------------------------------------------
// rand_chacha 0.3 `ChaCha20Rng::from_seed`: djb ChaCha20 with a 64-bit block counter and 64-bit stream id, a 64-word
// buffer (4 blocks); `next_u64` = lo | hi << 32 from two consecutive words.
class ChaCha20Rng {
   public:
    static ChaCha20Rng from_seed(const std::array<uint8_t, 32>& seed, uint64_t stream = 0) {
        ChaCha20Rng g;
        for (int i = 0; i < 8; i++) g.key_[i] = (uint32_t)seed[4 * i] | (uint32_t)seed[4 * i + 1] << 8 | (uint32_t)seed[4 * i + 2] << 16 | (uint32_t)seed[4 * i + 3] << 24;
        g.stream_ = stream;
        return g;
    }
    uint32_t next_u32() {
        if (idx_ >= 64) refill();
        return buf_[idx_++];
    }
    uint64_t next_u64() {
        const uint64_t lo = next_u32();
        const uint64_t hi = next_u32();
        return lo | hi << 32;
    }
    void fill_bytes(uint8_t* out, size_t n) {
        for (size_t i = 0; i < n; i += 4) {
            const uint32_t w = next_u32();
            for (size_t k = 0; k < 4 && i + k < n; k++) out[i + k] = (uint8_t)(w >> (8 * k));
        }
    }

   private:
    static uint32_t rotl(uint32_t v, int n) { return v << n | v >> (32 - n); }
    static void quarter(uint32_t* s, int a, int b, int c, int d) {
        s[a] += s[b]; s[d] = rotl(s[d] ^ s[a], 16);
        s[c] += s[d]; s[b] = rotl(s[b] ^ s[c], 12);
        s[a] += s[b]; s[d] = rotl(s[d] ^ s[a], 8);
        s[c] += s[d]; s[b] = rotl(s[b] ^ s[c], 7);
    }
    void block(uint32_t* out) {
        uint32_t init[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
        for (int i = 0; i < 8; i++) init[4 + i] = key_[i];
        init[12] = (uint32_t)counter_; init[13] = (uint32_t)(counter_ >> 32);
        init[14] = (uint32_t)stream_; init[15] = (uint32_t)(stream_ >> 32);
        uint32_t s[16];
        std::memcpy(s, init, sizeof(s));
        for (int r = 0; r < 10; r++) {
            quarter(s, 0, 4, 8, 12); quarter(s, 1, 5, 9, 13); quarter(s, 2, 6, 10, 14); quarter(s, 3, 7, 11, 15);
            quarter(s, 0, 5, 10, 15); quarter(s, 1, 6, 11, 12); quarter(s, 2, 7, 8, 13); quarter(s, 3, 4, 9, 14);
        }
        for (int i = 0; i < 16; i++) out[i] = s[i] + init[i];
        counter_++;
    }
    void refill() {
        for (int b = 0; b < 4; b++) block(buf_ + 16 * b);
        idx_ = 0;
    }
    uint32_t key_[8] = {}, buf_[64] = {};
    uint64_t stream_ = 0, counter_ = 0;
    int idx_ = 64;
};

// ark-ff 0.3 `Fp::rand` for BLS12-381 Fr (REPR_SHAVE_BITS = 1); returns the canonical value of the sampled element
template <class Rng>
Fr fr_rand(Rng& rng) {
    for (;;) {
        Fr v;
        for (int i = 0; i < 4; i++) v[i] = rng.next_u64();
        v[3] &= 0x7fffffffffffffffull;
        if (!detail::geq_mod(v.data())) return detail::from_montgomery(v);
    }
}

// ---- Proof ------------------------------------------------------------------------------------------------------------
class Proof {
   public:
    static constexpr size_t SIZE = MP_PROOF_BYTES;
    Proof() = default;
    // `TryFrom<Vec<u8>>` (groth16.rs:63-72): any other length is an Error
    static Result<Proof> try_from(const std::vector<uint8_t>& bytes) {
        if (bytes.size() != SIZE) return Error{};
        Proof p;
        std::memcpy(p.bytes_.data(), bytes.data(), SIZE);
        return p;
    }
    const std::array<uint8_t, SIZE>& to_bytes() const { return bytes_; }  // `proof_as_bytes`: compressed a | b | c
    // `codec::Encode`: the bytes as a Vec<u8>, u64-LE length prefix (manta-util/src/codec.rs:672-686)
    std::vector<uint8_t> encode() const {
        std::vector<uint8_t> out(8 + SIZE);
        const uint64_t n = SIZE;
        std::memcpy(out.data(), &n, 8);
        std::memcpy(out.data() + 8, bytes_.data(), SIZE);
        return out;
    }
    bool operator==(const Proof& o) const { return bytes_ == o.bytes_; }
    bool operator!=(const Proof& o) const { return !(*this == o); }
    std::array<uint8_t, SIZE>& raw() { return bytes_; }

   private:
    std::array<uint8_t, SIZE> bytes_{};
};

// ---- compiler -----------------------------------------------------------------------------------------------------------
// The per-circuit constant part of a finalized constraint system (ark `to_matrices()`): CSR, canonical coefficients,
// column index = instance variables first (0 is the constant 1), then witnesses.
struct SparseMatrix {
    std::vector<uint64_t> row_ptr;  // K + 1
    std::vector<uint32_t> col;
    std::vector<Fr> coeff;
};
struct ConstraintMatrices {
    uint64_t num_instance = 0, num_witness = 0;
    SparseMatrix a, b, c;
    uint64_t num_constraints() const { return a.row_ptr.empty() ? 0 : a.row_ptr.size() - 1; }
    uint64_t num_variables() const { return num_instance + num_witness; }
    mp_r1cs_view view() const {
        mp_r1cs_view v{};
        v.num_instance = num_instance;
        v.num_witness = num_witness;
        v.num_constraints = num_constraints();
        v.a_row_ptr = a.row_ptr.data(); v.a_col = a.col.data(); v.a_coeff = reinterpret_cast<const uint64_t*>(a.coeff.data());
        v.b_row_ptr = b.row_ptr.data(); v.b_col = b.col.data(); v.b_coeff = reinterpret_cast<const uint64_t*>(b.coeff.data());
        v.c_row_ptr = c.row_ptr.data(); v.c_col = c.col.data(); v.c_coeff = reinterpret_cast<const uint64_t*>(c.coeff.data());
        return v;
    }
};
// `R1CS<F>` as handed to `prove`: already synthesized; moved in and consumed by the call
struct R1CS {
    std::shared_ptr<const ConstraintMatrices> matrices;
    std::vector<Fr> assignment;  // z = [1, instance.., witness..], canonical
};

// ---- ProvingContext ---------------------------------------------------------------------------------------------------
// Owns the proving key in the reference's on-disk format (`serialize_unchecked`); device residency is created lazily per
// (device, circuit) and shared by copies, like the Arc the simulator's cloned contexts would share.
class ProvingContext {
   public:
    static Result<ProvingContext> decode(std::vector<uint8_t> bytes) {
        ProvingContext c;
        c.state_ = std::make_shared<State>();
        c.state_->bytes = std::move(bytes);
        if (mp_pk_parse(c.state_->bytes.data(), c.state_->bytes.size(), &c.state_->view) != MP_OK) return Error{};
        return c;
    }
    const std::vector<uint8_t>& encode() const { return state_->bytes; }
    bool operator==(const ProvingContext& o) const { return state_->bytes == o.state_->bytes; }

    // device handle for a circuit (nullptr on failure); keyed by the matrices object
    mp_ctx* native(const std::shared_ptr<const ConstraintMatrices>& m, int device) const {
        std::lock_guard<std::mutex> lock(state_->mu);
        auto key = std::make_pair(device, m.get());
        auto it = state_->handles.find(key);
        if (it != state_->handles.end()) return it->second.get();
        mp_ctx* h = nullptr;
        const mp_r1cs_view v = m->view();
        if (mp_ctx_create(&state_->view, &v, device, &h) != MP_OK) return nullptr;
        state_->handles.emplace(key, std::shared_ptr<mp_ctx>(h, [keep = m](mp_ctx* p) { mp_ctx_destroy(p); }));
        return h;
    }

   private:
    struct State {
        std::vector<uint8_t> bytes;
        mp_pk_view view{};
        std::mutex mu;
        std::map<std::pair<int, const ConstraintMatrices*>, std::shared_ptr<mp_ctx>> handles;
    };
    std::shared_ptr<State> state_;
};

// ---- the proof system ----------------------------------------------------------------------------------------------------
// `impl ProofSystem for Groth16<Bls12_381>` (groth16.rs:548-610), prove only: compile and verify stay on the reference's
// CPU path.  Every failure (malformed key, shape mismatch, CUDA error, no device) collapses to Error, like
// `.map_err(|_| Error)` at groth16.rs:597-599.
struct Groth16 {
    static inline int device = 0;

    template <class Rng>
    static Result<Proof> prove(const ProvingContext& context, R1CS compiler, Rng& rng) {
        const Fr r = fr_rand(rng);  // create_random_proof: r, then s, before anything else
        const Fr s = fr_rand(rng);
        return prove_with_randomness(context, std::move(compiler), r, s);
    }
    static Result<Proof> prove_with_randomness(const ProvingContext& context, R1CS compiler, const Fr& r, const Fr& s) {
        if (!compiler.matrices || compiler.assignment.size() != compiler.matrices->num_variables()) return Error{};
        mp_ctx* h = context.native(compiler.matrices, device);
        if (!h) return Error{};
        Proof p;
        if (mp_prove(h, reinterpret_cast<const uint64_t*>(compiler.assignment.data()), r.data(), s.data(), p.raw().data()) != MP_OK) return Error{};
        return p;
    }
    // Batch extension: the same proofs as looping `prove` with the same rng, in one device batch
    template <class Rng>
    static Result<std::vector<Proof>> prove_many(const ProvingContext& context, std::vector<R1CS> compilers, Rng& rng) {
        std::vector<Proof> out(compilers.size());
        if (compilers.empty()) return out;
        const auto& m = compilers[0].matrices;
        if (!m) return Error{};
        const size_t n = m->num_variables();
        std::vector<uint64_t> z(compilers.size() * n * 4), rs(compilers.size() * 4), ss(compilers.size() * 4);
        for (size_t i = 0; i < compilers.size(); i++) {
            if (compilers[i].matrices != m || compilers[i].assignment.size() != n) return Error{};
            const Fr r = fr_rand(rng), s = fr_rand(rng);
            std::memcpy(&rs[4 * i], r.data(), 32);
            std::memcpy(&ss[4 * i], s.data(), 32);
            std::memcpy(&z[i * n * 4], compilers[i].assignment.data(), n * 32);
        }
        mp_ctx* h = context.native(m, device);
        if (!h) return Error{};
        std::vector<uint8_t> bytes(compilers.size() * MP_PROOF_BYTES);
        if (mp_prove_batch(h, compilers.size(), z.data(), rs.data(), ss.data(), bytes.data()) != MP_OK) return Error{};
        for (size_t i = 0; i < compilers.size(); i++) std::memcpy(out[i].raw().data(), &bytes[i * MP_PROOF_BYTES], MP_PROOF_BYTES);
        return out;
    }
};

}  // namespace manta::groth16
