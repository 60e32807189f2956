// ntt_variant_bench.cu -- times ntt_tile_kernel<12, WHOLE> (forward + inverse) in isolation for kernel-variant
// experiments (compile with -DLSR_NTT_MINB=2|3|4 ...).  Links the product's host table builder.
// build: nvcc -O3 -std=c++17 -lineinfo -I lambda_snark_r_b200/csrc -I include -gencode arch=compute_100a,code=sm_100a \
//        tools/ntt_variant_bench.cu lambda_snark_r_b200/csrc/lsr_host.cpp -o tools/_bin/ntt_vb
#include <cstdio>
#include <vector>
#include "lsr_host.h"
#include "lsr_ntt.cuh"
#include <cstring>
using namespace lsr;
#ifndef VB_POL
#define VB_POL POL_F64
#endif
static void to_f64(std::vector<ulonglong2>& t, u64 q) {
    for (auto& e : t) { const double w = (double)e.x, wq = w / (double)q; memcpy(&e.x, &w, 8); memcpy(&e.y, &wq, 8); }
}
int main(int argc, char** argv) {
    const u64 q = 17592169062401ull; const uint32_t n = 4096; const size_t batch = argc > 1 ? atoi(argv[1]) : 16384;
    host::NttHostTables ht; host::build_ntt_tables(q, n, ht);
    ModParams mp = host::make_mod_params(q, 12);
    NttTables t{};
    if (VB_POL == POL_F64) {
        to_f64(ht.fwd, q); to_f64(ht.inv, q); to_f64(ht.fwd_last, q); to_f64(ht.inv_last, q);
        std::vector<ulonglong2> one{ht.n_inv}; to_f64(one, q); ht.n_inv = one[0];
    }
    ulonglong2 *df, *di, *dfl, *dil;
    cudaMalloc(&df, 16 * n); cudaMalloc(&di, 16 * n); cudaMalloc(&dfl, 16 * ht.fwd_last.size()); cudaMalloc(&dil, 16 * ht.inv_last.size());
    cudaMemcpy(df, ht.fwd.data(), 16 * n, cudaMemcpyHostToDevice); cudaMemcpy(di, ht.inv.data(), 16 * n, cudaMemcpyHostToDevice);
    cudaMemcpy(dfl, ht.fwd_last.data(), 16 * ht.fwd_last.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dil, ht.inv_last.data(), 16 * ht.inv_last.size(), cudaMemcpyHostToDevice);
    t.fwd = df; t.inv = di; t.fwd_last = dfl; t.inv_last = dil; t.n_inv = ht.n_inv;
    for (int i = 0; i < 16; i++) { t.head_fwd[i] = ht.fwd[i]; t.head_inv[i] = ht.inv[i]; }
    std::vector<u64> h(batch * n);
    for (size_t i = 0; i < h.size(); i++) h[i] = (i * 2654435761ull + 12345) % q;
    u64* d; cudaMalloc(&d, 8 * h.size()); cudaMemcpy(d, h.data(), 8 * h.size(), cudaMemcpyHostToDevice);
    auto kf = ntt_tile_kernel<12, true, VB_POL, false>; auto ki = ntt_tile_kernel<12, true, VB_POL, true>;
    constexpr size_t SMEM_F = 32768 + (ntt_pad<false>() ? 2048 : 0);      // padded layout of the forward kernel
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kf);
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kf, kNttThreads, 32768);
    printf("threads=%d POL=%d MINB=%d regs=%d occupancy=%d CTAs/SM\n", kNttThreads, VB_POL, LSR_NTT_MINB, fa.numRegs, occ);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int dir = 0; dir < 2; dir++) {
        float best = 1e9;
        for (int rep = 0; rep < 8; rep++) {
            cudaEventRecord(e0);
            if (dir == 0) kf<<<(unsigned)batch, kNttThreads, SMEM_F>>>(mp, t, d, batch * n, 0u, InvFusion{}); else ki<<<(unsigned)batch, kNttThreads, 32768>>>(mp, t, d, batch * n, 0u, InvFusion{});
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 1 && ms < best) best = ms;
        }
        printf("  %s: %.3f ms  %.2f M NTT/s\n", dir ? "inverse" : "forward", best, batch / best / 1e3);
    }
    std::vector<u64> back(h.size()); cudaMemcpy(back.data(), d, 8 * h.size(), cudaMemcpyDeviceToHost);
    size_t bad = 0; for (size_t i = 0; i < h.size(); i++) bad += back[i] != h[i];      // equal numbers of fwd and inv launches
    printf("  round trip mismatches: %zu\n", bad);
    return 0;
}
