// Host build of the transcript hash the device kernel runs (same source: csrc/lsr_keccak.h), so the CPU
// suite can pin it to hashlib.sha3_256 without a GPU.  Test infrastructure.
#include <cstdint>

#include "lsr_keccak.h"

extern "C" void lsr_test_fs_hash(const uint64_t* pub, uint64_t n_pub, const uint64_t* words, uint64_t n_words, uint64_t* out) {
    lsr::FsTranscript t{reinterpret_cast<const lsr::kw64*>(pub), n_pub, reinterpret_cast<const lsr::kw64*>(words), n_words};
    lsr::kw64 o[4];
    lsr::fs_sha3_256(t, o);
    for (int i = 0; i < 4; i++) out[i] = o[i];
}

// the lane-parallel formulation (index maps of fs_sha3_256_warp), emulated with arrays
extern "C" void lsr_test_fs_hash_lanes(const uint64_t* pub, uint64_t n_pub, const uint64_t* words, uint64_t n_words, uint64_t* out) {
    lsr::FsTranscript t{reinterpret_cast<const lsr::kw64*>(pub), n_pub, reinterpret_cast<const lsr::kw64*>(words), n_words};
    lsr::kw64 o[4];
    lsr::fs_sha3_256_lanes_emulated(t, o);
    for (int i = 0; i < 4; i++) out[i] = o[i];
}
