// lambda_snark_b200.hpp -- C++17 host-side mirror of the reference's safe wrappers over the C ABI.
//
// The reference's host side is Rust (rust-api/lambda-snark/src/{context,commitment,opening,challenge}.rs); there is no
// Rust toolchain in this image, so the same operator interface is written here in the other compiled language the image
// has, name for name and error for error, over exactly the symbols lambda-snark-sys binds.  Header-only; link
// liblambda_snark_core.  Exceptions stand in for Rust's Result::Err (CoreError::FfiError, CoreError::CommitmentFailed,
// Error::InvalidInput).
//
//   lsr::LweContext        context.rs:7-76        RAII over lwe_context_create / lwe_context_free, modulus()
//   lsr::Commitment        commitment.rs:14-110   new / linear_combine / clone / as_words, lwe_commitment_free on drop
//   lsr::Opening, generate_opening, verify_opening[_with_context]       opening.rs:20-264
//   lsr::Challenge::derive challenge.rs:102-134   through lsr_fs_challenge_batch (the transcript hash runs on the device)
//   lsr::NttContext        the ntt_* symbols (bound by bindgen, exercised by lambda-snark-sys/src/lib.rs:36-43)
#ifndef LAMBDA_SNARK_B200_HPP
#define LAMBDA_SNARK_B200_HPP

#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "lambda_snark_b200.h"

namespace lsr {

struct FfiError : std::runtime_error { using std::runtime_error::runtime_error; };            // CoreError::FfiError
struct CommitmentFailed : std::runtime_error { using std::runtime_error::runtime_error; };    // CoreError::CommitmentFailed
struct InvalidInput : std::invalid_argument { using std::invalid_argument::invalid_argument; };

// lambda-snark-core Params / Profile::RingB (lib.rs:129-196): validation rules of Params::validate
struct Params {
    uint32_t security_level = 128;
    uint32_t n = 4096, k = 2;
    uint64_t q = 17592186044417ull;
    double sigma = 3.19;
    void validate() const {
        if (n == 0 || (n & (n - 1))) throw InvalidInput("ring degree must be a power of two");
        if (k == 0) throw InvalidInput("module rank must be positive");
        if (q < (1ull << 24)) throw InvalidInput("modulus too small");
        if (sigma < 3.0) throw InvalidInput("sigma too small");
    }
};

class LweContext {                                              // context.rs:7-76
public:
    explicit LweContext(const Params& p) : modulus_(p.q) {
        p.validate();
        PublicParams c{};
        c.profile = PROFILE_RING_B; c.security_level = p.security_level; c.modulus = p.q;
        c.ring_degree = p.n; c.module_rank = p.k; c.sigma = p.sigma;
        inner_ = lwe_context_create(&c);
        if (!inner_) throw FfiError(std::string("lwe_context_create failed: ") + lsr_last_error());
    }
    LweContext(const LweContext&) = delete;
    LweContext& operator=(const LweContext&) = delete;
    LweContext(LweContext&& o) noexcept : inner_(std::exchange(o.inner_, nullptr)), modulus_(o.modulus_) {}
    ~LweContext() { lwe_context_free(inner_); }
    ::LweContext* as_ptr() const { return inner_; }
    uint64_t modulus() const { return modulus_; }               // the caller's field modulus, as context.rs:64-66
private:
    ::LweContext* inner_ = nullptr;
    uint64_t modulus_;
};

class Commitment {                                              // commitment.rs:14-110
public:
    // Commitment::new (commitment.rs:31-45): field elements reduced mod the context's modulus, then lwe_commit
    static Commitment create(const LweContext& ctx, const std::vector<uint64_t>& message, uint64_t seed) {
        std::vector<uint64_t> words(message.size());
        for (size_t i = 0; i < message.size(); i++) words[i] = message[i] % ctx.modulus();
        ::LweCommitment* c = lwe_commit(ctx.as_ptr(), words.data(), words.size(), seed);
        if (!c) throw CommitmentFailed(std::string("lwe_commit failed: ") + lsr_last_error());
        return Commitment(c);
    }
    // Commitment::linear_combine (commitment.rs:48-84)
    static Commitment linear_combine(const LweContext& ctx, const std::vector<const Commitment*>& commitments,
                                     const std::vector<uint64_t>& coeffs) {
        if (commitments.empty()) throw InvalidInput("no commitments provided");
        if (commitments.size() != coeffs.size()) throw InvalidInput("commitments/coeffs length mismatch");
        std::vector<const ::LweCommitment*> ptrs;
        for (const Commitment* c : commitments) ptrs.push_back(c->inner_);
        std::vector<uint64_t> words(coeffs.size());
        for (size_t i = 0; i < coeffs.size(); i++) words[i] = coeffs[i] % ctx.modulus();
        ::LweCommitment* r = lwe_linear_combine(ctx.as_ptr(), ptrs.data(), words.data(), ptrs.size());
        if (!r) throw CommitmentFailed(std::string("lwe_linear_combine failed: ") + lsr_last_error());
        return Commitment(r);
    }
    Commitment(const Commitment& o) : inner_(lwe_commitment_clone(o.inner_)) {          // impl Clone (commitment.rs:19-27)
        if (!inner_) throw FfiError("lwe_commitment_clone returned null");
    }
    Commitment(Commitment&& o) noexcept : inner_(std::exchange(o.inner_, nullptr)) {}
    Commitment& operator=(const Commitment&) = delete;
    ~Commitment() { lwe_commitment_free(inner_); }                                       // impl Drop
    // as_bytes (commitment.rs:87-93): the words the Fiat-Shamir transcript hashes; valid while the commitment lives
    const uint64_t* data() const { return inner_->data; }
    size_t len() const { return inner_->len; }
    const ::LweCommitment* as_ffi_ptr() const { return inner_; }
private:
    explicit Commitment(::LweCommitment* c) : inner_(c) {}
    ::LweCommitment* inner_;
};

inline uint64_t mul_mod(uint64_t a, uint64_t b, uint64_t q) { return (uint64_t)((unsigned __int128)a * b % q); }   // arith.rs:8-14
inline uint64_t add_mod(uint64_t a, uint64_t b, uint64_t q) { return (uint64_t)(((unsigned __int128)a + b) % q); }

// Polynomial::evaluate (polynomial.rs:97-113): Horner
inline uint64_t evaluate(const std::vector<uint64_t>& coeffs, uint64_t alpha, uint64_t q) {
    uint64_t acc = 0;
    for (size_t i = coeffs.size(); i-- > 0;) acc = add_mod(mul_mod(acc, alpha % q, q), coeffs[i] % q, q);
    return acc;
}

struct Opening {                                                // opening.rs:20-60
    uint64_t evaluation;
    std::vector<uint64_t> witness;                              // [randomness, coefficients...]
};

// generate_opening (opening.rs:104-115)
inline Opening generate_opening(const std::vector<uint64_t>& coeffs, uint64_t alpha, uint64_t randomness, uint64_t q) {
    Opening o{evaluate(coeffs, alpha, q), {randomness}};
    o.witness.insert(o.witness.end(), coeffs.begin(), coeffs.end());
    return o;
}

// verify_opening (opening.rs:229-264): evaluation consistency only
inline bool verify_opening(const Commitment&, uint64_t alpha, const Opening& o, uint64_t q) {
    if (o.evaluation >= q || o.witness.size() < 2) return false;
    std::vector<uint64_t> coeffs(o.witness.begin() + 1, o.witness.end());
    return evaluate(coeffs, alpha, q) == o.evaluation;
}

// verify_opening_with_context (opening.rs:160-222): + lwe_verify_opening of the commitment against the coefficients
inline bool verify_opening_with_context(const Commitment& c, uint64_t alpha, const Opening& o, uint64_t q,
                                        const LweContext& ctx) {
    if (!verify_opening(c, alpha, o, q)) return false;
    std::vector<uint64_t> msg(o.witness.size() - 1);
    for (size_t i = 0; i < msg.size(); i++) msg[i] = o.witness[i + 1] % q;
    uint64_t randomness = o.witness[0];
    LweOpening lo{&randomness, 1};
    return lwe_verify_opening(ctx.as_ptr(), c.as_ffi_ptr(), msg.data(), msg.size(), &lo) == 1;
}

struct Challenge {                                              // challenge.rs:20-134
    uint64_t alpha;
    std::array<uint8_t, 32> hash;
    // Challenge::derive: SHA3-256("LAMBDA-SNARK-R-FS-v1" || len || inputs || len || commitment words), alpha = LE64(h[0..8]) mod q
    static Challenge derive(const std::vector<uint64_t>& public_inputs, const Commitment& c, uint64_t q) {
        uint64_t ab[2] = {0, 0}, h[8] = {0};
        if (lsr_fs_challenge_batch(public_inputs.data(), public_inputs.size(), c.data(), c.len(), 1, q, 0, ab, h) != 0)
            throw FfiError(std::string("lsr_fs_challenge_batch failed: ") + lsr_last_error());
        Challenge r{ab[0], {}};
        for (int i = 0; i < 32; i++) r.hash[(size_t)i] = (uint8_t)(h[i / 8] >> (8 * (i % 8)));
        return r;
    }
};

class NttContext {                                              // ntt.h:24-92 behind an RAII handle
public:
    NttContext(uint64_t q, uint32_t n) : inner_(ntt_context_create(q, n)), n_(n) {
        if (!inner_) throw FfiError(std::string("ntt_context_create failed: ") + lsr_last_error());
    }
    NttContext(const NttContext&) = delete;
    NttContext& operator=(const NttContext&) = delete;
    ~NttContext() { ntt_context_free(inner_); }
    void forward(std::vector<uint64_t>& coeffs) const {
        if (ntt_forward(inner_, coeffs.data(), (uint32_t)coeffs.size()) != 0) throw InvalidInput("ntt_forward: bad arguments");
    }
    void inverse(std::vector<uint64_t>& evals) const {
        if (ntt_inverse(inner_, evals.data(), (uint32_t)evals.size()) != 0) throw InvalidInput("ntt_inverse: bad arguments");
    }
    std::vector<uint64_t> mul_pointwise(const std::vector<uint64_t>& a, const std::vector<uint64_t>& b) const {
        if (a.size() != n_ || b.size() != n_) throw InvalidInput("ntt_mul_pointwise: wrong length");
        std::vector<uint64_t> r(n_);
        ntt_mul_pointwise(inner_, r.data(), a.data(), b.data(), n_);
        return r;
    }
private:
    ::NttContext* inner_;
    uint32_t n_;
};

}  // namespace lsr
#endif  // LAMBDA_SNARK_B200_HPP
