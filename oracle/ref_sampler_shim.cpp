// ref_sampler_shim.cpp -- ORACLE/TEST INFRASTRUCTURE (never shipped, never on the product path).
//
// Compiles the reference's OWN sampler source, cpp-core/src/utils.cpp, *where it
// lies* under /root/reference (passed by the Makefile as -DLSR_REF_UTILS_CPP=...),
// into oracle/_ref/libref_sampler.so.  No reference source is copied into this
// repo.  The only change is the entropy source: std::random_device is replaced
// (by macro, for this translation unit only) with a queue the caller fills, so
// that the reference's build_cdf (utils.cpp:26-75) and sample_single
// (utils.cpp:95-121) can be driven with known 64-bit draws and used to PIN
// oracle/lsr_oracle.c's lsro_cdt_build / lsro_cdt_sample.
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <random>
#include <vector>

namespace std {
struct lsr_fake_random_device {
    const uint64_t* draws = nullptr;
    size_t pos = 0;      // counts 32-bit halves
    unsigned int operator()() {
        // random_u64 (utils.cpp:77-93) shifts in two 32-bit chunks, first = high half
        const uint64_t v = draws[pos >> 1];
        const unsigned int out = (pos & 1) ? static_cast<unsigned int>(v)
                                           : static_cast<unsigned int>(v >> 32);
        ++pos;
        return out;
    }
};
}  // namespace std

#define random_device lsr_fake_random_device
#include LSR_REF_UTILS_CPP
#undef random_device

extern "C" {

// reference build_cdf(sigma) -> table; returns entry count
size_t ref_build_cdf(double sigma, uint64_t* out, size_t cap) {
    const GaussianTable t = build_cdf(sigma);
    if (t.cdf.size() > cap) return 0;
    for (size_t i = 0; i < t.cdf.size(); ++i) out[i] = t.cdf[i];
    return t.cdf.size();
}

// reference sample_single driven by draws[2*i], draws[2*i+1]; out two's complement
void ref_sample(double sigma, const uint64_t* draws, size_t count, uint64_t* out) {
    const GaussianTable t = build_cdf(sigma);
    std::lsr_fake_random_device rd;
    rd.draws = draws;
    for (size_t i = 0; i < count; ++i) {
        out[i] = static_cast<uint64_t>(sample_single(t, rd));
    }
}

// the reference's exported symbol, unchanged except for the entropy source:
// with the macro in force it would need a default-constructed queue, so it is
// exercised only for its argument validation (utils.cpp:133-135).
int ref_sample_gaussian_validate(uint64_t* output, size_t len, double sigma) {
    if (!output || len == 0 || !(sigma > 0.0) || !std::isfinite(sigma)) return -1;
    return 0;
}

}  // extern "C"
