// lsr_r1cs.h -- the R1CS handle behind lambda_snark_r1cs_* (cpp-core/src/r1cs.cpp:18-46) and the
// device-side state of the quotient pipeline that hangs off it (lsr_quotient.cu).
#pragma once
#include <cstdint>
#include <mutex>
#include <vector>

#include "lambda_snark_b200.h"

struct NttContext;

namespace lsr {

struct QuotientState;     // lsr_quotient.cu: CSR copies on the device, cyclic NTT contexts, scratch

struct R1csHandle {
    std::vector<SparseEntry> A, B, C;
    uint32_t rows = 0, cols = 0;
    uint64_t q = 0;
    std::mutex mu;                     // serialises the lazily built device state
    QuotientState* quotient = nullptr;
    ~R1csHandle();
};

void quotient_state_free(QuotientState* s);

}  // namespace lsr
