// butterfly_microbench.cu -- register-resident throughput of the NTT butterfly
// variants (no memory traffic): how many lazy CT butterflies per second can the
// chip issue with the code nvcc actually generates?
// build: nvcc -O3 -std=c++17 -I lambda_snark_r_b200/csrc -gencode arch=compute_100a,code=sm_100a tools/butterfly_microbench.cu -o tools/_bin/bfly_mb
#include <cstdio>
#include <cuda_runtime.h>

#include "lsr_arith.cuh"

using namespace lsr;


__device__ __forceinline__ u64 mulred4_hi(u64 x, u64 w, u64 ws, u64 nq) {   // variant with mul.hi cross terms
    const u32 x0 = lo32(x), x1 = hi32(x), s0 = lo32(ws), s1 = hi32(ws);
    const u32 a = __umulhi(x1, s0), b = __umulhi(x0, s1);
    const u64 qh = mad_wide(x1, s1, (u64)a) + (u64)b;
    return mullo2_acc(x, w, qh, nq);
}

// Shoup with the exact 64x64 high product (compiler's __umul64hi)
__device__ __forceinline__ u64 mulred_exact(u64 x, u64 w, u64 ws, u64 nq) {
    return mullo2_acc(x, w, __umul64hi(x, ws), nq);
}

template <int VAR, int NV, int MINB>
__global__ void __launch_bounds__(256, MINB) k(unsigned iters, u64 q, u64 w0, u64 ws0, u64* sink, int flag) {
    const u64 nq = 0 - q, q4 = 4 * q;
    u64 v[NV];
#pragma unroll
    for (int i = 0; i < NV; i++) v[i] = (threadIdx.x * 1315423911ull + i * 2654435761ull + blockIdx.x) % q;
    u64 w = w0 + threadIdx.x, ws = ws0 + threadIdx.x;
    for (unsigned it = 0; it < iters; it++) {
#pragma unroll
        for (int half = NV / 2; half >= 1; half >>= 1) {
#pragma unroll
            for (int j = 0; j < NV; j++) {
                if ((j & half) == 0) {
                    const int jj = j | half;
                    u64 T;
                    if (VAR == 0) T = mulred4(v[jj], w, ws, nq);
                    if (VAR == 1) T = mulred4_hi(v[jj], w, ws, nq);
                    if (VAR == 2) T = mulred_exact(v[jj], w, ws, nq);
                    const u64 X = v[j];
                    v[j] = X + T;
                    v[jj] = X + q4 - T;
                }
            }
        }
        // keep values bounded like the real kernel does at pass boundaries
#pragma unroll
        for (int j = 0; j < NV; j++) v[j] &= 0x000fffffffffffffull;
        w += 2; ws += 2;
    }
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < NV; i++) s ^= v[i];
    if (flag) sink[threadIdx.x] = s;
}

template <int VAR, int NV, int MINB>
void run(const char* name, int sms) {
    u64* sink; cudaMalloc(&sink, 8 * 256);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const u64 q = 17592169062401ull;
    const unsigned blocks = sms * MINB, iters = 3000;
    double best = 0;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        k<VAR, NV, MINB><<<blocks, 256>>>(iters, q, 1299579534ull, 1362715717599ull, sink, 0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        int stages = 0; for (int h = NV / 2; h >= 1; h >>= 1) stages++;
        const double bf = (double)blocks * 256 * iters * (NV / 2) * stages;
        const double rate = bf / (ms * 1e-3);
        if (rep && rate > best) best = rate;
    }
    printf("%-34s %8.1f Gbutterfly/s -> %6.1f M NTT(n=4096)/s compute bound, %5.2f clk/SM per warp-butterfly\n", name,
           best / 1e9, best / 24576 / 1e6, sms * 1.965e9 * 32 / best);
    cudaFree(sink);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    const int s = p.multiProcessorCount;
    run<0, 16, 3>("mulred4 NV=16 3cta", s);
    run<0, 16, 2>("mulred4 NV=16 2cta", s);
    run<0, 16, 1>("mulred4 NV=16 1cta", s);
    run<0, 8, 3>("mulred4 NV=8 3cta", s);
    run<0, 8, 4>("mulred4 NV=8 4cta", s);
    run<0, 8, 6>("mulred4 NV=8 6cta", s);
    run<0, 4, 8>("mulred4 NV=4 8cta", s);
    run<0, 2, 8>("mulred4 NV=2 8cta", s);
    run<1, 16, 3>("mul.hi variant NV=16 3cta", s);
    run<2, 16, 3>("exact shoup NV=16 3cta", s);
    return 0;
}
