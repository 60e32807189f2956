// lsr_r1cs.cpp -- the five lambda_snark_r1cs_* symbols that
// rust-api/lambda-snark-core/src/r1cs.rs:121-141 declares by hand and that must
// exist at link time once SEAL/NTL are gone.
//
// Host-only: the sparse mat-vecs are a few thousand multiply-adds; there is no
// data-parallel work worth a kernel launch.  Semantics follow
// cpp-core/src/r1cs.cpp:18-180 and cpp-core/src/ffi.cpp:27-105 without NTL:
// values pass through static_cast<long> before reduction (r1cs.cpp:165-167), so
// a stored u64 of 2^64-1 means -1 mod q.
#include <cstdio>
#include <cstring>
#include <iomanip>
#include <sstream>
#include <string>
#include <new>
#include <stdexcept>
#include <vector>

#include "lambda_snark_b200.h"
#include "lsr_r1cs.h"

namespace lsr {
R1csHandle::~R1csHandle() { quotient_state_free(quotient); }
}  // namespace lsr

namespace {

typedef unsigned __int128 u128;
using lsr::R1csHandle;

// NTL::conv(ZZ_p, long): signed value reduced into [0, q)
inline uint64_t to_field(uint64_t raw, uint64_t q) {
    const int64_t sv = static_cast<int64_t>(raw);
    if (sv >= 0) return static_cast<uint64_t>(sv) % q;
    const uint64_t mag = (0 - static_cast<uint64_t>(sv)) % q;
    return mag ? q - mag : 0;
}

inline uint64_t mulmod(uint64_t a, uint64_t b, uint64_t q) { return (uint64_t)(((u128)a * b) % q); }

// r1cs.cpp:148-180
std::vector<uint64_t> sparse_mv(const std::vector<SparseEntry>& m, uint32_t rows,
                                const std::vector<uint64_t>& z, uint64_t q) {
    std::vector<uint64_t> out(rows, 0);
    for (const SparseEntry& e : m) {
        if (e.col >= z.size()) throw std::out_of_range("R1CS: Column index exceeds witness length");
        if (e.row >= rows) throw std::out_of_range("R1CS: Row index exceeds constraint count");
        const uint64_t term = mulmod(to_field(e.value, q), to_field(z[e.col], q), q);
        uint64_t acc = to_field(out[e.row], q) + term;
        if (acc >= q) acc -= q;
        out[e.row] = acc;
    }
    return out;
}

}  // namespace

extern "C" {

LambdaSnarkError lambda_snark_r1cs_create(const SparseMatrix* A, const SparseMatrix* B,
                                          const SparseMatrix* C, uint64_t modulus, void** out_r1cs) {
    if (!A || !B || !C || !out_r1cs) return LAMBDA_SNARK_ERR_NULL_PTR;            // ffi.cpp:34-36
    try {
        if (A->n_rows != B->n_rows || B->n_rows != C->n_rows ||
            A->n_cols != B->n_cols || B->n_cols != C->n_cols || modulus < 2) {
            return LAMBDA_SNARK_ERR_INVALID_PARAMS;                                // r1cs.cpp:23-28
        }
        R1csHandle* h = new R1csHandle;
        h->rows = A->n_rows; h->cols = A->n_cols; h->q = modulus;
        auto copy = [](const SparseMatrix* m, std::vector<SparseEntry>& dst) {    // deep copy, r1cs.cpp:34-46
            if (m->n_entries && m->entries) dst.assign(m->entries, m->entries + m->n_entries);
        };
        copy(A, h->A); copy(B, h->B); copy(C, h->C);
        *out_r1cs = h;
        return LAMBDA_SNARK_OK;
    } catch (const std::bad_alloc&) {
        return LAMBDA_SNARK_ERR_ALLOC_FAILED;
    } catch (...) {
        return LAMBDA_SNARK_ERR_CRYPTO_FAILED;
    }
}

LambdaSnarkError lambda_snark_r1cs_validate_witness(void* r1cs, const R1CSWitness* witness, bool* out_valid) {
    if (!r1cs || !witness || !out_valid) return LAMBDA_SNARK_ERR_NULL_PTR;         // ffi.cpp:62-64
    try {
        const R1csHandle* h = static_cast<const R1csHandle*>(r1cs);
        if (witness->len != h->cols) return LAMBDA_SNARK_ERR_INVALID_PARAMS;       // r1cs.cpp:100-105
        if (witness->len == 0 || !witness->values) return LAMBDA_SNARK_ERR_INVALID_PARAMS;
        std::vector<uint64_t> z(witness->values, witness->values + witness->len);
        if (z[0] != 1) return LAMBDA_SNARK_ERR_INVALID_PARAMS;                     // r1cs.cpp:108-110
        const std::vector<uint64_t> az = sparse_mv(h->A, h->rows, z, h->q);
        const std::vector<uint64_t> bz = sparse_mv(h->B, h->rows, z, h->q);
        const std::vector<uint64_t> cz = sparse_mv(h->C, h->rows, z, h->q);
        bool ok = true;
        for (uint32_t i = 0; i < h->rows && ok; ++i) ok = mulmod(az[i], bz[i], h->q) == cz[i];
        *out_valid = ok;
        return LAMBDA_SNARK_OK;
    } catch (const std::invalid_argument&) {
        return LAMBDA_SNARK_ERR_INVALID_PARAMS;
    } catch (...) {
        return LAMBDA_SNARK_ERR_CRYPTO_FAILED;                                     // out_of_range lands here, ffi.cpp:73-75
    }
}

void lambda_snark_r1cs_free(void* r1cs) {
    delete static_cast<R1csHandle*>(r1cs);
}

uint32_t lambda_snark_r1cs_num_constraints(void* r1cs) {
    return r1cs ? static_cast<R1csHandle*>(r1cs)->rows : 0;
}

uint32_t lambda_snark_r1cs_num_variables(void* r1cs) {
    return r1cs ? static_cast<R1csHandle*>(r1cs)->cols : 0;
}


/* ---- Lean 4 export: cpp-core/src/lean_ffi.cpp:152-314.  Pure host formatting; the two
 * SEAL-specific exports have nothing to serialise in this library (no SEAL context, no SEAL public key)
 * and fail the way the reference does when its SEAL context is missing (:251-254, :293-296).        */
namespace {
std::string sparse_matrix_to_lean(const SparseMatrix& m) {               // lean_ffi.cpp:46-61
    std::ostringstream oss;
    oss << "SparseMatrix.mk " << m.n_rows << " " << m.n_cols << " [";
    for (size_t i = 0; i < m.n_entries; ++i) {
        if (i > 0) oss << ", ";
        oss << "(" << m.entries[i].row << ", " << m.entries[i].col << ", " << m.entries[i].value << ")";
    }
    oss << "]";
    return oss.str();
}
int copy_out(const std::string& text, char* out, size_t cap, const char* who) {
    if (text.size() + 1 > cap) {
        std::fprintf(stderr, "%s: buffer too small (need %zu, have %zu)\n", who, text.size() + 1, cap);
        return -1;
    }
    std::memcpy(out, text.c_str(), text.size() + 1);
    return static_cast<int>(text.size());
}
}  // namespace

int export_vk_to_lean(const R1CSConstraintSystem* r1cs, const PublicParams* params, char* out_buffer,
                      size_t buffer_size) LSR_NOEXCEPT {
    if (!r1cs || !params || !out_buffer) return -1;
    try {
        if (r1cs->n_public_inputs > r1cs->n_vars) {                      // lean_ffi.cpp:161-167
            std::fprintf(stderr, "export_vk_to_lean: n_public_inputs (%u) exceeds n_vars (%u)\n",
                         r1cs->n_public_inputs, r1cs->n_vars);
            return -1;
        }
        for (const SparseMatrix* m : {&r1cs->A, &r1cs->B, &r1cs->C})
            if (m->n_entries && !m->entries) return -1;
        std::ostringstream oss;
        oss << "\xE2\x9F\xA8" << r1cs->n_constraints << ", " << r1cs->n_vars << ", " << r1cs->n_public_inputs << ", "
            << params->modulus << ", " << sparse_matrix_to_lean(r1cs->A) << ", " << sparse_matrix_to_lean(r1cs->B)
            << ", " << sparse_matrix_to_lean(r1cs->C) << "\xE2\x9F\xA9";   // U+27E8 ... U+27E9
        return copy_out(oss.str(), out_buffer, buffer_size, "export_vk_to_lean");
    } catch (...) {
        return -1;
    }
}

int export_params_to_lean(const PublicParams* params, char* out_buffer, size_t buffer_size) LSR_NOEXCEPT {
    if (!params || !out_buffer) return -1;
    try {
        std::ostringstream oss;                                          // lean_ffi.cpp:68-78
        oss << std::fixed << std::setprecision(1);
        oss << "{ n := " << params->ring_degree << ", k := " << params->module_rank << ", q := " << params->modulus
            << ", \xCF\x83 := " << params->sigma << ", \xCE\xBB := " << params->security_level << " }";
        const std::string text = oss.str();
        if (text.size() + 1 > buffer_size) return -1;
        std::memcpy(out_buffer, text.c_str(), text.size() + 1);
        return static_cast<int>(text.size());
    } catch (...) {
        return -1;
    }
}

int export_seal_context_to_lean(const LweContext* ctx, char* out_buffer, size_t) LSR_NOEXCEPT {
    if (!ctx || !out_buffer) return -1;
    std::fprintf(stderr, "export_seal_context_to_lean: SEAL context not initialized\n");
    return -1;
}

int export_seal_pubkey_to_lean(const LweContext* ctx, char* out_buffer, size_t) LSR_NOEXCEPT {
    if (!ctx || !out_buffer) return -1;
    std::fprintf(stderr, "export_seal_pubkey_to_lean: SEAL context or key not initialized\n");
    return -1;
}

}  // extern "C"
