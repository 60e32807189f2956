// lsr_microbench.cu -- measures the integer-multiply roofline denominator.
//
// MEASURED_PEAKS.json records HBM and tensor peaks only; the NTT is bound by
// the 32-bit IMAD pipe, so its peak is measured here the same way the driver
// measures the others: a dependency-free loop of mad.wide.u32 / mad.lo.u32 on
// every SM, timed with CUDA events ("the builder must measure it").
#include <cuda_runtime.h>

#include "lambda_snark_b200.h"
#include "lsr_engine.h"

namespace lsr {

constexpr int kChains = 8;     // independent accumulators per thread (latency 4, issue every 2)
constexpr int kInner = 64;

// Every multiplicand is produced by the previous instruction of a neighbouring
// chain, so ptxas can neither hoist a product out of the loop nor split the
// multiply-add (it does both with loop-invariant operands: an earlier version of
// this kernel ended up timing 64-bit adds).  WIDE: IMAD.WIDE d64 = a*b (full
// 64-bit product, the form the NTT butterflies use); else IMAD d = a*b + d.
template <bool WIDE>
__global__ void __launch_bounds__(256)
imad_peak_kernel(unsigned iters, unsigned seed, unsigned long long* sink, int flag) {
    unsigned b[kChains], d[kChains];
    unsigned long long D[kChains];
#pragma unroll
    for (int c = 0; c < kChains; c++) {
        b[c] = (blockIdx.x * 40503u + threadIdx.x * 2654435761u + c * 1315423911u + seed) | 1u;
        d[c] = b[c] * 7u + c;
        D[c] = ((unsigned long long)d[c] << 32) | b[c];
    }
    for (unsigned it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < kInner; r++) {
#pragma unroll
            for (int c = 0; c < kChains; c++) {
                const int n = (c + 1) % kChains;
                if (WIDE) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(D[c]) : "r"((unsigned)(D[n] >> 32)), "r"(b[c]));
                else      asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(d[c]) : "r"(d[n]), "r"(b[c]));
            }
        }
    }
    unsigned long long s = 0;
#pragma unroll
    for (int c = 0; c < kChains; c++) s += D[c] + d[c];
    if (flag) sink[threadIdx.x] = s;
}

// FP64 pipe (the POL_F64 butterflies live on it): chains of fma.rn.f64 whose multiplicand is the
// previous result of a neighbouring chain.  DADD and DMUL issue at the same rate as DFMA
// (tools/fp64_microbench.cu), so one number serves as the denominator for all three.
__global__ void __launch_bounds__(256)
fp64_peak_kernel(unsigned iters, unsigned seed, double* sink, int flag) {
    double a[kChains], b[kChains];
#pragma unroll
    for (int c = 0; c < kChains; c++) {
        a[c] = 1.0 + 1e-9 * (double)(threadIdx.x + c + seed);
        b[c] = 1.0 - 1e-9 * (double)(blockIdx.x + c);
    }
    for (unsigned it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < kInner; r++) {
#pragma unroll
            for (int c = 0; c < kChains; c++) {
                const int n = (c + 1) % kChains;
                asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a[c]) : "d"(a[n]), "d"(b[c]));
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < kChains; c++) s += a[c];
    if (flag) sink[threadIdx.x] = s;
}

}  // namespace lsr

extern "C" int lsr_measure_fp64_peak(double* ginst_per_s) LSR_NOEXCEPT {
    using namespace lsr;
    if (!ginst_per_s) return -1;
    int dev = current_device_choice();
    if (!cuda_ok(cudaSetDevice(dev), "cudaSetDevice")) return -1;
    cudaDeviceProp prop;
    if (!cuda_ok(cudaGetDeviceProperties(&prop, dev), "cudaGetDeviceProperties")) return -1;
    double* sink = nullptr;
    if (!cuda_ok(cudaMalloc(&sink, 256 * sizeof(*sink)), "cudaMalloc")) return -1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const unsigned blocks = prop.multiProcessorCount * 8;
    const unsigned iters = 1000;
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        fp64_peak_kernel<<<blocks, 256>>>(iters, rep, sink, 0);
        cudaEventRecord(e1);
        if (!cuda_ok(cudaEventSynchronize(e1), "fp64_peak_kernel")) { cudaFree(sink); return -1; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double rate = (double)blocks * 256.0 * iters * kInner * kChains / (ms * 1e-3) / 1e9;
        if (rep > 0 && rate > best) best = rate;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    *ginst_per_s = best;
    return 0;
}

extern "C" int lsr_measure_imad_peak(int wide, double* gimad_per_s, double* sm_mhz_effective) LSR_NOEXCEPT {
    using namespace lsr;
    if (!gimad_per_s) return -1;
    int dev = current_device_choice();
    if (!cuda_ok(cudaSetDevice(dev), "cudaSetDevice")) return -1;
    cudaDeviceProp prop;
    if (!cuda_ok(cudaGetDeviceProperties(&prop, dev), "cudaGetDeviceProperties")) return -1;
    unsigned long long* sink = nullptr;
    if (!cuda_ok(cudaMalloc(&sink, 256 * sizeof(*sink)), "cudaMalloc")) return -1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const unsigned blocks = prop.multiProcessorCount * 8;    // 8 CTAs x 256 threads = 64 warps / SM
    const unsigned iters = 2000;
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        if (wide) imad_peak_kernel<true><<<blocks, 256>>>(iters, rep, sink, 0);
        else      imad_peak_kernel<false><<<blocks, 256>>>(iters, rep, sink, 0);
        cudaEventRecord(e1);
        if (!cuda_ok(cudaEventSynchronize(e1), "imad_peak_kernel")) { cudaFree(sink); return -1; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double ops = (double)blocks * 256.0 * iters * kInner * kChains;
        const double rate = ops / (ms * 1e-3) / 1e9;
        if (rep > 0 && rate > best) best = rate;     // first repetition is warm-up
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    *gimad_per_s = best;
    if (sm_mhz_effective) *sm_mhz_effective = best * 1e3 / (64.0 * prop.multiProcessorCount);   // if 64 lanes/clk/SM
    return 0;
}
