// Parity test of the C++ host mirror (include/mantaprover.hpp), written like the reference's own prove tests
// (manta-pay/src/test/transfer.rs:62-109: sample a context, build the compiler, `prove(&context, compiler, &mut rng)`),
// except that the expected proof bytes are known: the fixture is written by tests/test_cpp_host.py (CPU oracle).
//
//   host_mirror_test <fixture>   exit 0: every proof matches; 3: prove returned Error; 1: mismatch / bad fixture
#include <cstdio>
#include <fstream>
#include <iterator>

#include "mantaprover.hpp"

using namespace manta::groth16;

struct Reader {
    std::vector<uint8_t> d;
    size_t pos = 0;
    bool ok = true;
    uint64_t u64() {
        uint64_t v = 0;
        raw(&v, 8);
        return v;
    }
    void raw(void* out, size_t n) {
        if (pos + n > d.size()) { ok = false; std::memset(out, 0, n); return; }
        std::memcpy(out, d.data() + pos, n);
        pos += n;
    }
};

static SparseMatrix read_matrix(Reader& r, uint64_t K) {
    SparseMatrix m;
    const uint64_t nnz = r.u64();
    m.row_ptr.resize(K + 1);
    m.col.resize(nnz);
    m.coeff.resize(nnz);
    r.raw(m.row_ptr.data(), (K + 1) * 8);
    r.raw(m.col.data(), nnz * 4);
    r.raw(m.coeff.data(), nnz * 32);
    return m;
}

int main(int argc, char** argv) {
    if (argc < 2) return 1;
    std::ifstream f(argv[1], std::ios::binary);
    Reader r;
    r.d.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
    std::vector<uint8_t> pk(r.u64());
    r.raw(pk.data(), pk.size());
    auto mats = std::make_shared<ConstraintMatrices>();
    mats->num_instance = r.u64();
    mats->num_witness = r.u64();
    const uint64_t K = r.u64();
    mats->a = read_matrix(r, K);
    mats->b = read_matrix(r, K);
    mats->c = read_matrix(r, K);
    const uint64_t count = r.u64(), n = mats->num_variables();
    std::vector<std::vector<Fr>> zs(count, std::vector<Fr>(n));
    std::vector<std::array<uint8_t, 32>> seeds(count);
    std::vector<std::array<uint8_t, 192>> expect(count), expect_many(count);
    for (uint64_t i = 0; i < count; i++) {
        r.raw(zs[i].data(), n * 32);
        r.raw(seeds[i].data(), 32);
        r.raw(expect[i].data(), 192);
        r.raw(expect_many[i].data(), 192);
    }
    if (!r.ok || r.pos != r.d.size()) { std::fprintf(stderr, "bad fixture\n"); return 1; }

    // a malformed key is an Error, not a crash
    if (!ProvingContext::decode(std::vector<uint8_t>(pk.begin(), pk.begin() + 100)).is_err()) return 1;
    if (!Proof::try_from(std::vector<uint8_t>(191)).is_err()) return 1;

    auto ctx_r = ProvingContext::decode(pk);
    if (ctx_r.is_err()) { std::fprintf(stderr, "decode failed\n"); return 1; }
    const ProvingContext& context = ctx_r.unwrap();
    int bad = 0;
    for (uint64_t i = 0; i < count; i++) {
        auto rng = ChaCha20Rng::from_seed(seeds[i]);
        auto proof = Groth16::prove(context, R1CS{mats, zs[i]}, rng);
        if (proof.is_err()) { std::printf("prove -> Error\n"); return 3; }
        if (proof.unwrap().to_bytes() != expect[i]) { std::printf("proof %llu differs\n", (unsigned long long)i); bad++; }
        auto enc = proof.unwrap().encode();
        if (enc.size() != 200 || enc[0] != 192) bad++;
    }
    // prove_many == a loop over prove with one rng
    {
        auto rng = ChaCha20Rng::from_seed(seeds[0]);
        std::vector<R1CS> compilers;
        for (uint64_t i = 0; i < count; i++) compilers.push_back(R1CS{mats, zs[i]});
        auto proofs = Groth16::prove_many(context, std::move(compilers), rng);
        if (proofs.is_err()) { std::printf("prove_many -> Error\n"); return 3; }
        for (uint64_t i = 0; i < count; i++)
            if (proofs.unwrap()[i].to_bytes() != expect_many[i]) { std::printf("batch proof %llu differs\n", (unsigned long long)i); bad++; }
    }
    // a wrong-length assignment collapses to Error as well
    {
        auto rng = ChaCha20Rng::from_seed(seeds[0]);
        std::vector<Fr> shorter(zs[0].begin(), zs[0].end() - 1);
        if (!Groth16::prove(context, R1CS{mats, shorter}, rng).is_err()) bad++;
    }
    std::printf(bad ? "MISMATCH\n" : "ok %llu proofs bit-exact\n", (unsigned long long)count);
    return bad ? 1 : 0;
}
