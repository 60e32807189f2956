// lsr_ntt.cu -- NttContext life cycle, kernel instantiation and launchers for
// K1 (forward), K2 (inverse), K3 (pointwise).  Replaces the SEAL calls of the
// reference's cpp-core/src/ntt.cpp.
#include <algorithm>
#include <cstdio>
#include <cstring>

#include <cstdlib>

#include "lsr_engine.h"
#include "lsr_ntt.cuh"

namespace lsr {

// ------------------------------------------------------------------ utilities
static thread_local std::string g_error;
static thread_local int g_device_choice = 0;

void set_error(const std::string& msg) { g_error = msg; }
const char* last_error() { return g_error.c_str(); }
int current_device_choice() { return g_device_choice; }
void set_device_choice(int dev) { g_device_choice = dev; }

bool cuda_ok(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return true;
    std::string msg = std::string(what) + ": " + cudaGetErrorString(e);
    set_error(msg);
    std::fprintf(stderr, "lambda_snark_b200: %s\n", msg.c_str());   // reference logs to stderr too
    return false;
}

bool DeviceScratch::reserve(size_t need) {
    if (need <= bytes) return true;
    if (ptr) cudaFree(ptr);
    ptr = nullptr; bytes = 0;
    if (!cuda_ok(cudaMalloc(&ptr, need), "cudaMalloc(scratch)")) { ptr = nullptr; return false; }
    bytes = need;
    return true;
}
void DeviceScratch::release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr; bytes = 0;
}
bool PinnedScratch::reserve(size_t need) {
    if (need <= bytes) return true;
    if (ptr) cudaFreeHost(ptr);
    ptr = nullptr; bytes = 0;
    if (!cuda_ok(cudaMallocHost(&ptr, need), "cudaMallocHost")) { ptr = nullptr; return false; }
    bytes = need;
    return true;
}
void PinnedScratch::release() {
    if (ptr) cudaFreeHost(ptr);
    ptr = nullptr; bytes = 0;
}

// ------------------------------------------------------------------- launches
template <typename K>
static bool ensure_smem(K kernel, size_t smem) {
    if (smem <= 48 * 1024) return true;
    return cuda_ok(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                   "cudaFuncSetAttribute(smem)");
}

// Arithmetic policy of a transform: FP64 butterflies whenever they are exact for the modulus
// (q < 2^45, unless the context was pinned to integer arithmetic), else the u64 paths.
static int policy_of(const NttContext* c, bool inverse) {
    if (c->mp.gold) return POL_GOLD;
    if (c->arith != 1 && c->mp.f64_ok) return POL_F64;
    return (inverse ? c->mp.lazy_inv : c->mp.lazy_fwd) ? POL_LAZY : POL_GUARD;
}

// d = log n - LT for block tiles (WHOLE = false), ignored otherwise
template <int LT, bool WHOLE, bool INV, bool FUSED = false>
static bool launch_tile(const NttContext* c, u64* d, size_t total, cudaStream_t s, u32 dlog = 0, const InvFusion fz = InvFusion{}) {
    constexpr int TL = LT > kTileLogMin ? LT : kTileLogMin;
    const size_t smem = (sizeof(u64) << TL) + (ntt_pad<INV, LT>() ? (sizeof(u64) << (TL - 4)) : 0);
    const size_t tiles = (total + ((size_t)1 << TL) - 1) >> TL;
    if (tiles == 0) return true;
    if (tiles > 0x7fffffffull) { set_error("batch too large for one launch"); return false; }
    const int pol = policy_of(c, INV);
    if (pol == POL_F64) {
        auto k = ntt_tile_kernel<LT, WHOLE, POL_F64, INV, FUSED>;
        if (!ensure_smem(k, smem)) return false;
        k<<<(unsigned)tiles, kNttThreads, smem, s>>>(c->mp, c->tables_f, d, total, dlog, fz);
    } else if (pol == POL_LAZY) {
        auto k = ntt_tile_kernel<LT, WHOLE, POL_LAZY, INV, FUSED>;
        if (!ensure_smem(k, smem)) return false;
        k<<<(unsigned)tiles, kNttThreads, smem, s>>>(c->mp, c->tables, d, total, dlog, fz);
    } else if (pol == POL_GOLD) {
        auto k = ntt_tile_kernel<LT, WHOLE, POL_GOLD, INV, FUSED>;
        if (!ensure_smem(k, smem)) return false;
        k<<<(unsigned)tiles, kNttThreads, smem, s>>>(c->mp, c->tables, d, total, dlog, fz);
    } else {
        auto k = ntt_tile_kernel<LT, WHOLE, POL_GUARD, INV, FUSED>;
        if (!ensure_smem(k, smem)) return false;
        k<<<(unsigned)tiles, kNttThreads, smem, s>>>(c->mp, c->tables, d, total, dlog, fz);
    }
    return cuda_ok(cudaGetLastError(), "ntt_tile_kernel launch");
}

template <int S, bool INV, bool FIRST>
static bool launch_column(const NttContext* c, u64* d, size_t batch, cudaStream_t s, u32 logn, u32 s0,
                          const u64* fin_c = nullptr, u64 fin_scale = 0) {
    const size_t cols = batch << (logn - S);
    const size_t blocks = (cols + kNttThreads - 1) / kNttThreads;
    if (blocks == 0) return true;
    if (blocks > 0x7fffffffull) { set_error("batch too large for one launch"); return false; }
    const int pol = policy_of(c, INV);
    if constexpr (INV && FIRST) {
        if (fin_c) {
            if (pol == POL_F64)       ntt_column_kernel<S, POL_F64, true, true, true><<<(unsigned)blocks, kNttThreads, 0, s>>>(c->mp, c->tables_f, d, batch, logn, s0, fin_c, fin_scale);
            else if (pol == POL_LAZY) ntt_column_kernel<S, POL_LAZY, true, true, true><<<(unsigned)blocks, kNttThreads, 0, s>>>(c->mp, c->tables, d, batch, logn, s0, fin_c, fin_scale);
            else if (pol == POL_GOLD) ntt_column_kernel<S, POL_GOLD, true, true, true><<<(unsigned)blocks, kNttThreads, 0, s>>>(c->mp, c->tables, d, batch, logn, s0, fin_c, fin_scale);
            else                      ntt_column_kernel<S, POL_GUARD, true, true, true><<<(unsigned)blocks, kNttThreads, 0, s>>>(c->mp, c->tables, d, batch, logn, s0, fin_c, fin_scale);
            return cuda_ok(cudaGetLastError(), "ntt_column_kernel launch");
        }
    }
    if (pol == POL_F64)       ntt_column_kernel<S, POL_F64, INV, FIRST><<<(unsigned)blocks, kNttThreads, 0, s>>>(c->mp, c->tables_f, d, batch, logn, s0, fin_c, fin_scale);
    else if (pol == POL_LAZY) ntt_column_kernel<S, POL_LAZY, INV, FIRST><<<(unsigned)blocks, kNttThreads, 0, s>>>(c->mp, c->tables, d, batch, logn, s0, fin_c, fin_scale);
    else if (pol == POL_GOLD) ntt_column_kernel<S, POL_GOLD, INV, FIRST><<<(unsigned)blocks, kNttThreads, 0, s>>>(c->mp, c->tables, d, batch, logn, s0, fin_c, fin_scale);
    else                      ntt_column_kernel<S, POL_GUARD, INV, FIRST><<<(unsigned)blocks, kNttThreads, 0, s>>>(c->mp, c->tables, d, batch, logn, s0, fin_c, fin_scale);
    return cuda_ok(cudaGetLastError(), "ntt_column_kernel launch");
}

template <bool INV, bool FIRST>
static bool launch_column_s(int S, const NttContext* c, u64* d, size_t batch, cudaStream_t s, u32 logn, u32 s0,
                            const u64* fin_c = nullptr, u64 fin_scale = 0) {
    switch (S) {
        case 1: return launch_column<1, INV, FIRST>(c, d, batch, s, logn, s0, fin_c, fin_scale);
        case 2: return launch_column<2, INV, FIRST>(c, d, batch, s, logn, s0, fin_c, fin_scale);
        case 3: return launch_column<3, INV, FIRST>(c, d, batch, s, logn, s0, fin_c, fin_scale);
        case 4: return launch_column<4, INV, FIRST>(c, d, batch, s, logn, s0, fin_c, fin_scale);
        case 5: return launch_column<5, INV, FIRST>(c, d, batch, s, logn, s0, fin_c, fin_scale);
        default: set_error("bad column pass"); return false;
    }
}

// S1 + S2 column stages in one HBM round trip (ntt_column2_kernel)
template <int S1, int S2, bool INV, bool FIRST>
static bool launch_column2(const NttContext* c, u64* d, size_t batch, cudaStream_t s, u32 logn, u32 s0,
                           const u64* fin_c = nullptr, u64 fin_scale = 0) {
    const size_t blocks = batch << (logn - 12);          // 4096 coefficients per CTA
    if (blocks == 0) return true;
    if (blocks > 0x7fffffffull) { set_error("batch too large for one launch"); return false; }
    const int pol = policy_of(c, INV);
    if constexpr (INV && FIRST) {
        if (fin_c) {
            if (pol == POL_F64)       ntt_column2_kernel<S1, S2, POL_F64, true, true, true><<<(unsigned)blocks, kNttThreads, 0, s>>>(c->mp, c->tables_f, d, batch, logn, s0, fin_c, fin_scale);
            else if (pol == POL_LAZY) ntt_column2_kernel<S1, S2, POL_LAZY, true, true, true><<<(unsigned)blocks, kNttThreads, 0, s>>>(c->mp, c->tables, d, batch, logn, s0, fin_c, fin_scale);
            else if (pol == POL_GOLD) ntt_column2_kernel<S1, S2, POL_GOLD, true, true, true><<<(unsigned)blocks, kNttThreads, 0, s>>>(c->mp, c->tables, d, batch, logn, s0, fin_c, fin_scale);
            else                      ntt_column2_kernel<S1, S2, POL_GUARD, true, true, true><<<(unsigned)blocks, kNttThreads, 0, s>>>(c->mp, c->tables, d, batch, logn, s0, fin_c, fin_scale);
            return cuda_ok(cudaGetLastError(), "ntt_column2_kernel launch");
        }
    }
    if (pol == POL_F64)       ntt_column2_kernel<S1, S2, POL_F64, INV, FIRST><<<(unsigned)blocks, kNttThreads, 0, s>>>(c->mp, c->tables_f, d, batch, logn, s0, fin_c, fin_scale);
    else if (pol == POL_LAZY) ntt_column2_kernel<S1, S2, POL_LAZY, INV, FIRST><<<(unsigned)blocks, kNttThreads, 0, s>>>(c->mp, c->tables, d, batch, logn, s0, fin_c, fin_scale);
    else if (pol == POL_GOLD) ntt_column2_kernel<S1, S2, POL_GOLD, INV, FIRST><<<(unsigned)blocks, kNttThreads, 0, s>>>(c->mp, c->tables, d, batch, logn, s0, fin_c, fin_scale);
    else                      ntt_column2_kernel<S1, S2, POL_GUARD, INV, FIRST><<<(unsigned)blocks, kNttThreads, 0, s>>>(c->mp, c->tables, d, batch, logn, s0, fin_c, fin_scale);
    return cuda_ok(cudaGetLastError(), "ntt_column2_kernel launch");
}

static bool column2_enabled() {
#ifdef LSR_PROFILING    // tools/ build only: LSR_NTT_COLUMN2=0 falls back to chained single-column passes
    static const bool on = [] { const char* e = std::getenv("LSR_NTT_COLUMN2"); return !(e && e[0] == '0'); }();
    return on;
#else
    return true;
#endif
}

// n <= 2^13: one kernel, the polynomial never leaves shared memory.
// n >= 2^14: column passes (first log n - 12 stages, at most 5 per pass: one pass up to 2^17, two up to
// 2^22, three up to 2^24) + tile kernel on 4096-blocks.
// (n = 2^14 fits one CTA's shared memory, but at one 8-warp CTA per SM it ran at
// 625 G butterflies/s against 800 for the two-kernel path: profiles/r01_ntt_sweep.json)
// (Issuing the two kernels in L2-sized batch chunks so the second finds the first one's output on chip
// was measured and rejected: 16/32/64 MB chunks ran at 0.60/0.77/0.85 of the whole-batch rate -- the
// small dependent launches cost more in tails than the second HBM round trip does.)
template <int LOGN, bool INV, bool FUSED>
static bool launch_whole(const NttContext* c, u64* d, size_t batch, cudaStream_t s, const InvFusion& fz) {
    return launch_tile<LOGN, true, INV, FUSED>(c, d, batch << LOGN, s, 0u, fz);
}

// fz (inverse only): the first kernel (tile kernel on 4096-blocks) takes the product and the destination, the kernel that
// ends the transform takes the finishing step; everything after the first kernel works in place on the destination
template <bool INV, bool FUSED>
static bool launch_big(const NttContext* c, u64* d_in, size_t batch, cudaStream_t s, const InvFusion& fz) {
    const InvFusion first{fz.mul, fz.dst, nullptr, 0};
    u64* d = (INV && fz.dst) ? fz.dst : d_in;
    const u64* fin_c = INV ? fz.fin_c : nullptr;
    const u64 fin_scale = fz.fin_scale;
    const u32 logn = c->logn;
    const int C = (int)logn - 12;                       // column stages
    const int passes = (C + 4) / 5;
    int S[3] = {0, 0, 0}, s0[3] = {0, 0, 0};
    for (int i = 0, left = C, at = 0; i < passes; i++) {
        S[i] = (left + (passes - i) - 1) / (passes - i);   // as even as possible, larger passes first
        s0[i] = at;
        at += S[i];
        left -= S[i];
    }
    const size_t total = batch << logn;
    if (C >= 6 && column2_enabled()) {
        // 6 .. 8 column stages: one shared-memory pass; 9 .. 12: a register pass of C - 8 stages + one of 8
        const int head = C > 8 ? C - 8 : 0;
        auto fused = [&](bool first) -> bool {
            if (C == 6) return launch_column2<3, 3, INV, true>(c, d, batch, s, logn, 0u, fin_c, fin_scale);
            if (C == 7) return launch_column2<4, 3, INV, true>(c, d, batch, s, logn, 0u, fin_c, fin_scale);
            return first ? launch_column2<4, 4, INV, true>(c, d, batch, s, logn, 0u, fin_c, fin_scale)
                         : launch_column2<4, 4, INV, false>(c, d, batch, s, logn, (u32)head);
        };
        if (!INV) {
            if (head && !launch_column_s<false, true>(head, c, d, batch, s, logn, 0u)) return false;
            if (!fused(head == 0)) return false;
            return launch_tile<12, false, false>(c, d, total, s, (u32)C);
        }
        if (!launch_tile<12, false, true, FUSED>(c, d_in, total, s, (u32)C, first)) return false;
        if (!fused(head == 0)) return false;
        return head ? launch_column_s<true, true>(head, c, d, batch, s, logn, 0u, fin_c, fin_scale) : true;
    }
    if (!INV) {
        for (int i = 0; i < passes; i++) {
            const bool ok = i == 0 ? launch_column_s<false, true>(S[i], c, d, batch, s, logn, (u32)s0[i])
                                   : launch_column_s<false, false>(S[i], c, d, batch, s, logn, (u32)s0[i]);
            if (!ok) return false;
        }
        return launch_tile<12, false, false>(c, d, total, s, (u32)C);
    }
    if (!launch_tile<12, false, true, FUSED>(c, d_in, total, s, (u32)C, first)) return false;
    for (int i = passes - 1; i >= 0; i--) {
        const bool ok = i == 0 ? launch_column_s<true, true>(S[i], c, d, batch, s, logn, (u32)s0[i], fin_c, fin_scale)
                               : launch_column_s<true, false>(S[i], c, d, batch, s, logn, (u32)s0[i]);
        if (!ok) return false;
    }
    return true;
}

static bool plan_is_one_pass(uint32_t logn);

template <bool INV, bool FUSED = false>
static bool dispatch(const NttContext* c, u64* d, size_t batch, cudaStream_t s, const InvFusion fz = InvFusion{}) {
    if ((reinterpret_cast<uintptr_t>(d) | reinterpret_cast<uintptr_t>(fz.dst)) & 15u) {     // unit-stride passes use 16-byte global accesses
        set_error("device polynomial buffer must be 16-byte aligned");
        return false;
    }
    if ((fz.mul || fz.dst || fz.fin_c) && (!FUSED || plan_is_one_pass(c->logn))) {
        set_error("fused inverse transform: unsupported shape");           // callers fall back to separate kernels
        return false;
    }
    switch (c->logn) {
        case 1: return launch_whole<1, INV, FUSED>(c, d, batch, s, fz);
        case 2: return launch_whole<2, INV, FUSED>(c, d, batch, s, fz);
        case 3: return launch_whole<3, INV, FUSED>(c, d, batch, s, fz);
        case 4: return launch_whole<4, INV, FUSED>(c, d, batch, s, fz);
        case 5: return launch_whole<5, INV, FUSED>(c, d, batch, s, fz);
        case 6: return launch_whole<6, INV, FUSED>(c, d, batch, s, fz);
        case 7: return launch_whole<7, INV, FUSED>(c, d, batch, s, fz);
        case 8: return launch_whole<8, INV, FUSED>(c, d, batch, s, fz);
        case 9: return launch_whole<9, INV, FUSED>(c, d, batch, s, fz);
        case 10: return launch_whole<10, INV, FUSED>(c, d, batch, s, fz);
        case 11: return launch_whole<11, INV, FUSED>(c, d, batch, s, fz);
        case 12: return launch_whole<12, INV, FUSED>(c, d, batch, s, fz);
        case 13: return launch_whole<13, INV, FUSED>(c, d, batch, s, fz);
        default:
            if (c->logn >= 14 && c->logn <= (uint32_t)kMaxEngineLogN) return launch_big<INV, FUSED>(c, d, batch, s, fz);
            set_error("unsupported ring degree");
            return false;
    }
}

// one-pass plans (n <= 16) read and write global memory inside the single pass: no staging loop to fuse into
static bool plan_is_one_pass(uint32_t logn) { return logn <= 4; }

bool ntt_inverse_fused_supported(const NttContext* ctx) { return !plan_is_one_pass(ctx->logn); }

bool ntt_inverse_fused_launch(const NttContext* ctx, u64* d_data, size_t batch, cudaStream_t stream, const InvFusion& fz) {
    return dispatch<true, true>(ctx, d_data, batch, stream, fz);
}

bool ntt_forward_launch(const NttContext* ctx, u64* d_data, size_t batch, cudaStream_t stream) {
    return dispatch<false>(ctx, d_data, batch, stream);
}
bool ntt_inverse_launch(const NttContext* ctx, u64* d_data, size_t batch, cudaStream_t stream) {
    return dispatch<true>(ctx, d_data, batch, stream);
}

// in-place bit-reversal permutation of every polynomial of the batch (index i <-> brv_logn(i))
static __global__ void __launch_bounds__(256)
bitrev_permute_kernel(u64* __restrict__ data, int logn, size_t total) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const u32 i = (u32)(idx & (((size_t)1 << logn) - 1));
    const u32 j = __brev(i) >> (32 - logn);
    if (i < j) {
        u64* p = data + (idx - i);
        const u64 a = p[i], b = p[j];
        p[i] = b; p[j] = a;
    }
}

bool ntt_bitrev_launch(const NttContext* ctx, u64* d_data, size_t batch, cudaStream_t stream) {
    const size_t total = batch << ctx->logn;
    if (total == 0) return true;
    const size_t blocks = (total + 255) / 256;
    if (blocks > 0x7fffffffull) { set_error("batch too large for one launch"); return false; }
    bitrev_permute_kernel<<<(unsigned)blocks, 256, 0, stream>>>(d_data, (int)ctx->logn, total);
    return cuda_ok(cudaGetLastError(), "bitrev_permute_kernel launch");
}

bool pointwise_launch(const NttContext* ctx, u64* d_r, const u64* d_a, const u64* d_b, size_t total,
                      cudaStream_t stream) {
    if (total == 0) return true;
    size_t blocks = (total / 2 + 256 * kPointwiseU - 1) / (256 * kPointwiseU);
    blocks = std::max<size_t>(1, std::min<size_t>(blocks, 148 * 32));   // tile-stride loop, multiple of the SM count
    if (ctx->mp.gold) pointwise_mul_kernel<true><<<(unsigned)blocks, 256, 0, stream>>>(ctx->mp, d_r, d_a, d_b, total);
    else pointwise_mul_kernel<false><<<(unsigned)blocks, 256, 0, stream>>>(ctx->mp, d_r, d_a, d_b, total);
    return cuda_ok(cudaGetLastError(), "pointwise_mul_kernel launch");
}

// ------------------------------------------------------------------ life cycle
static ulonglong2 to_f64_pair(const ulonglong2& e, u64 q) {
    const double w = (double)e.x, wq = w / (double)q;      // w < 2^45: exact; IEEE division
    ulonglong2 r;
    std::memcpy(&r.x, &w, 8);
    std::memcpy(&r.y, &wq, 8);
    return r;
}
static std::vector<ulonglong2> to_f64_pairs(const std::vector<ulonglong2>& t, u64 q) {
    std::vector<ulonglong2> r(t.size());
    for (size_t i = 0; i < t.size(); ++i) r[i] = to_f64_pair(t[i], q);
    return r;
}

static NttContext* ntt_create_from(const host::NttHostTables& ht, bool cyclic);

NttContext* ntt_create(u64 q, uint32_t n) {
    host::NttHostTables ht;
    if (!host::build_ntt_tables(q, n, ht)) {
        set_error("ntt_context_create: invalid (q, n)");
        return nullptr;
    }
    return ntt_create_from(ht, false);
}

NttContext* ntt_create_negacyclic(u64 q, uint32_t n, u64 psi) {
    host::NttHostTables ht;
    if ((q != kGoldilocks && !host::is_prime(q)) || !host::build_negacyclic_tables(q, n, psi, ht)) {
        set_error("negacyclic NTT context: invalid (q, n, psi)");
        return nullptr;
    }
    return ntt_create_from(ht, false);
}

NttContext* ntt_create_cyclic(u64 q, uint32_t n, u64 omega) {
    host::NttHostTables ht;
    if (!host::build_cyclic_tables(q, n, omega, ht)) {
        set_error("cyclic NTT context: invalid (q, n, omega)");
        return nullptr;
    }
    return ntt_create_from(ht, true);
}

static NttContext* ntt_create_from(const host::NttHostTables& ht, bool cyclic) {
    const u64 q = ht.q;
    const uint32_t n = ht.n;
    const int dev = current_device_choice();
    if (!cuda_ok(cudaSetDevice(dev), "cudaSetDevice")) return nullptr;
    NttContext* c = new (std::nothrow) NttContext;
    if (!c) return nullptr;
    c->modulus = q; c->degree = n; c->logn = ht.logn; c->psi = ht.psi; c->device = dev; c->cyclic = cyclic;
    c->mp = host::make_mod_params(q, ht.logn);
    const size_t bytes = sizeof(ulonglong2) * n;
    bool ok = cuda_ok(cudaMalloc(&c->d_fwd, bytes), "cudaMalloc(twiddles)") &&
              cuda_ok(cudaMalloc(&c->d_inv, bytes), "cudaMalloc(twiddles)") &&
              cuda_ok(cudaMemcpy(c->d_fwd, ht.fwd.data(), bytes, cudaMemcpyHostToDevice), "upload twiddles") &&
              cuda_ok(cudaMemcpy(c->d_inv, ht.inv.data(), bytes, cudaMemcpyHostToDevice), "upload twiddles") &&
              cuda_ok(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking), "cudaStreamCreate");
    if (ok && !ht.fwd_last.empty()) {
        const size_t lb = sizeof(ulonglong2) * ht.fwd_last.size();
        ok = cuda_ok(cudaMalloc(&c->d_fwd_last, lb), "cudaMalloc(twiddles)") &&
             cuda_ok(cudaMalloc(&c->d_inv_last, lb), "cudaMalloc(twiddles)") &&
             cuda_ok(cudaMemcpy(c->d_fwd_last, ht.fwd_last.data(), lb, cudaMemcpyHostToDevice), "upload twiddles") &&
             cuda_ok(cudaMemcpy(c->d_inv_last, ht.inv_last.data(), lb, cudaMemcpyHostToDevice), "upload twiddles");
    }
    for (int i = 0; ok && i < 2; ++i)
        ok = cuda_ok(cudaStreamCreateWithFlags(&c->copy_streams[i], cudaStreamNonBlocking), "cudaStreamCreate");
    // POL_F64 tables: the same entries as (double)w, RN(w / q) bit patterns
    std::vector<ulonglong2> f_fwd, f_inv, f_fwd_last, f_inv_last;
    if (ok && c->mp.f64_ok) {
        f_fwd = to_f64_pairs(ht.fwd, q); f_inv = to_f64_pairs(ht.inv, q);
        f_fwd_last = to_f64_pairs(ht.fwd_last, q); f_inv_last = to_f64_pairs(ht.inv_last, q);
        ok = cuda_ok(cudaMalloc(&c->d_f_fwd, bytes), "cudaMalloc(twiddles)") &&
             cuda_ok(cudaMalloc(&c->d_f_inv, bytes), "cudaMalloc(twiddles)") &&
             cuda_ok(cudaMemcpy(c->d_f_fwd, f_fwd.data(), bytes, cudaMemcpyHostToDevice), "upload twiddles") &&
             cuda_ok(cudaMemcpy(c->d_f_inv, f_inv.data(), bytes, cudaMemcpyHostToDevice), "upload twiddles");
        if (ok && !f_fwd_last.empty()) {
            const size_t lb = sizeof(ulonglong2) * f_fwd_last.size();
            ok = cuda_ok(cudaMalloc(&c->d_f_fwd_last, lb), "cudaMalloc(twiddles)") &&
                 cuda_ok(cudaMalloc(&c->d_f_inv_last, lb), "cudaMalloc(twiddles)") &&
                 cuda_ok(cudaMemcpy(c->d_f_fwd_last, f_fwd_last.data(), lb, cudaMemcpyHostToDevice), "upload twiddles") &&
                 cuda_ok(cudaMemcpy(c->d_f_inv_last, f_inv_last.data(), lb, cudaMemcpyHostToDevice), "upload twiddles");
        }
    }
    if (!ok) { ntt_destroy(c); return nullptr; }
    if (c->mp.f64_ok) {
        c->tables_f.fwd = c->d_f_fwd;
        c->tables_f.inv = c->d_f_inv;
        c->tables_f.fwd_last = c->d_f_fwd_last;
        c->tables_f.inv_last = c->d_f_inv_last;
        for (uint32_t i = 0; i < 32; ++i) {
            c->tables_f.head_fwd[i] = i < n ? f_fwd[i] : ulonglong2{0, 0};
            c->tables_f.head_inv[i] = i < n ? f_inv[i] : ulonglong2{0, 0};
        }
        c->tables_f.n_inv = to_f64_pair(ht.n_inv, q);
    }
    c->tables.fwd = c->d_fwd;
    c->tables.inv = c->d_inv;
    for (uint32_t i = 0; i < 32; ++i) {
        c->tables.head_fwd[i] = i < n ? ht.fwd[i] : ulonglong2{0, 0};
        c->tables.head_inv[i] = i < n ? ht.inv[i] : ulonglong2{0, 0};
    }
    c->tables.fwd_last = c->d_fwd_last;
    c->tables.inv_last = c->d_inv_last;
    c->tables.n_inv = ht.n_inv;
    return c;
}

void ntt_destroy(NttContext* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (auto& s : c->scratch) s.release();
    c->pin.release();
    if (c->d_fwd) cudaFree(c->d_fwd);
    if (c->d_inv) cudaFree(c->d_inv);
    if (c->d_fwd_last) cudaFree(c->d_fwd_last);
    if (c->d_inv_last) cudaFree(c->d_inv_last);
    if (c->d_f_fwd) cudaFree(c->d_f_fwd);
    if (c->d_f_inv) cudaFree(c->d_f_inv);
    if (c->d_f_fwd_last) cudaFree(c->d_f_fwd_last);
    if (c->d_f_inv_last) cudaFree(c->d_f_inv_last);
    if (c->stream) cudaStreamDestroy(c->stream);
    for (auto& s : c->copy_streams) if (s) cudaStreamDestroy(s);
    delete c;
}

// ---------------------------------------------------------------- host paths
// Chunks of <= 32 MiB alternate between two streams, each running
// H2D -> kernel -> D2H in order, so copies of one chunk overlap the kernel of
// the other (fully so when the caller's memory is page-locked).
bool ntt_transform_host(const NttContext* c, u64* host, size_t batch, bool inverse, bool natural) {
    if (batch == 0) return true;
    std::lock_guard<std::mutex> lock(c->mu);
    if (!cuda_ok(cudaSetDevice(c->device), "cudaSetDevice")) return false;
    const size_t n = c->degree;
    const size_t poly_bytes = n * sizeof(u64);
    // Small single calls (the reference's one-polynomial ntt_forward / ntt_inverse): no device buffer and no copy
    // commands at all.  The polynomial is copied into a page-locked buffer, the one kernel of the transform (n <= 8192)
    // works on that buffer in place through its device mapping -- 32 KiB each way over PCIe inside the kernel -- and the
    // result is copied back: one launch + one synchronisation instead of two staged pageable copies around them.
    if (batch * poly_bytes <= ((size_t)256 << 10) && c->logn <= 13 && !natural) {
        if (!c->pin.reserve((size_t)256 << 10)) return false;
        u64* buf = static_cast<u64*>(c->pin.ptr);
        std::memcpy(buf, host, batch * poly_bytes);
        const bool ok = (inverse ? ntt_inverse_launch(c, buf, batch, c->stream) : ntt_forward_launch(c, buf, batch, c->stream)) &&
                        cuda_ok(cudaStreamSynchronize(c->stream), "sync");
        if (ok) std::memcpy(host, buf, batch * poly_bytes);
        return ok;
    }
    size_t chunk = std::max<size_t>(1, ((size_t)32 << 20) / poly_bytes);
    chunk = std::min(chunk, batch);
    const int nbuf = batch > chunk ? 2 : 1;
    for (int b = 0; b < nbuf; ++b)
        if (!c->scratch[b].reserve(chunk * poly_bytes)) return false;
    size_t done = 0;
    int b = 0;
    bool ok = true;
    while (ok && done < batch) {
        const size_t cnt = std::min(chunk, batch - done);
        cudaStream_t s = nbuf == 1 ? c->stream : c->copy_streams[b];
        u64* d = static_cast<u64*>(c->scratch[b].ptr);
        u64* h = host + done * n;
        ok = cuda_ok(cudaMemcpyAsync(d, h, cnt * poly_bytes, cudaMemcpyHostToDevice, s), "H2D") &&
             (inverse ? ((!natural || ntt_bitrev_launch(c, d, cnt, s)) && ntt_inverse_launch(c, d, cnt, s))
                      : (ntt_forward_launch(c, d, cnt, s) && (!natural || ntt_bitrev_launch(c, d, cnt, s)))) &&
             cuda_ok(cudaMemcpyAsync(h, d, cnt * poly_bytes, cudaMemcpyDeviceToHost, s), "D2H");
        done += cnt;
        b ^= 1;
    }
    if (nbuf == 1) ok = cuda_ok(cudaStreamSynchronize(c->stream), "sync") && ok;
    else for (int i = 0; i < 2; ++i) ok = cuda_ok(cudaStreamSynchronize(c->copy_streams[i]), "sync") && ok;
    return ok;
}

bool pointwise_host(const NttContext* c, u64* r, const u64* a, const u64* b, size_t total) {
    if (total == 0) return true;
    std::lock_guard<std::mutex> lock(c->mu);
    if (!cuda_ok(cudaSetDevice(c->device), "cudaSetDevice")) return false;
    const size_t chunk = std::min<size_t>(total, (size_t)4 << 20);   // 32 MiB per operand
    for (int i = 0; i < 2; ++i)
        if (!c->scratch[i].reserve(chunk * sizeof(u64))) return false;
    u64* da = static_cast<u64*>(c->scratch[0].ptr);
    u64* db = static_cast<u64*>(c->scratch[1].ptr);
    cudaStream_t s = c->stream;
    bool ok = true;
    for (size_t done = 0; ok && done < total; done += chunk) {
        const size_t cnt = std::min(chunk, total - done);
        ok = cuda_ok(cudaMemcpyAsync(da, a + done, cnt * sizeof(u64), cudaMemcpyHostToDevice, s), "H2D") &&
             cuda_ok(cudaMemcpyAsync(db, b + done, cnt * sizeof(u64), cudaMemcpyHostToDevice, s), "H2D") &&
             pointwise_launch(c, da, da, db, cnt, s) &&
             cuda_ok(cudaMemcpyAsync(r + done, da, cnt * sizeof(u64), cudaMemcpyDeviceToHost, s), "D2H") &&
             cuda_ok(cudaStreamSynchronize(s), "sync");
    }
    return ok;
}

}  // namespace lsr
