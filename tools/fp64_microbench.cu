// fp64_microbench.cu -- is the FP64 pipe of B200 a usable second multiplier for the NTT?
//   1. DFMA / DADD / DMUL issue rate (lanes/clk/SM)
//   2. does DFMA overlap with IMAD and with ALU work issued from the same warps?
//   3. register-resident throughput of the FP64 butterfly (exact modmul by fma error-free
//      product + rounded quotient, values held as doubles)
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/fp64_microbench.cu -o tools/_bin/fp64_mb
#include <cstdio>
#include <cuda_runtime.h>

constexpr int CH = 8, INNER = 32;

template <int OP>
__global__ void __launch_bounds__(256) k(unsigned iters, unsigned seed, double* sink, int flag) {
    double a[CH], b[CH];
    unsigned d[CH], e[CH];
#pragma unroll
    for (int c = 0; c < CH; c++) {
        a[c] = 1.0 + 1e-9 * (threadIdx.x + c + seed);
        b[c] = 1.0 - 1e-9 * (blockIdx.x + c);
        d[c] = threadIdx.x * 2654435761u + c + seed;
        e[c] = d[c] * 7u + 1u;
    }
    for (unsigned it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < INNER; r++) {
#pragma unroll
            for (int c = 0; c < CH; c++) {
                const int n = (c + 1) % CH;
                if (OP == 0) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a[c]) : "d"(a[n]), "d"(b[c]));
                if (OP == 1) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(a[c]) : "d"(b[n]));
                if (OP == 2) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(a[c]) : "d"(b[n]));
                if (OP == 3) {  // DFMA + independent IMAD.lo
                    asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a[c]) : "d"(a[n]), "d"(b[c]));
                    asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(d[c]) : "r"(d[n]), "r"(e[c]));
                }
                if (OP == 4) {  // DFMA + independent LOP3
                    asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a[c]) : "d"(a[n]), "d"(b[c]));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(d[c]) : "r"(d[n]), "r"(e[c]));
                }
                if (OP == 5) {  // DFMA + IMAD.lo + LOP3
                    asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a[c]) : "d"(a[n]), "d"(b[c]));
                    asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(d[c]) : "r"(d[n]), "r"(e[c]));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(e[c]) : "r"(e[n]), "r"(d[n]));
                }
                if (OP == 6) {  // IMAD.lo + LOP3 (reference point)
                    asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(d[c]) : "r"(d[n]), "r"(e[c]));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(e[c]) : "r"(e[n]), "r"(d[n]));
                }
                if (OP == 7) {  // 2 DFMA + 1 IMAD.WIDE
                    asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(a[c]) : "d"(a[n]), "d"(b[c]));
                    asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(b[c]) : "d"(b[n]), "d"(a[n]));
                    unsigned long long D = ((unsigned long long)e[c] << 32) | d[c];
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(D) : "r"(d[n]), "r"(e[n]));
                    d[c] = (unsigned)D; e[c] = (unsigned)(D >> 32);
                }
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CH; c++) s += a[c] + b[c] + d[c] + e[c];
    if (flag) sink[threadIdx.x] = s;
}

template <int OP> void run(const char* name, int sms, double mhz) {
    double* sink; cudaMalloc(&sink, 8 * 256);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const unsigned blocks = sms * 8, iters = 2000; double best = 0;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0); k<OP><<<blocks, 256>>>(iters, rep, sink, 0); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double rate = (double)blocks * 256 * iters * INNER * CH / (ms * 1e-3);
        if (rep && rate > best) best = rate;
    }
    printf("%-40s %7.2f op-groups/clk/SM (%5.2f clk per warp-group per SMSP)\n", name, best / (sms * mhz * 1e6),
           4 * 32 / (best / (sms * mhz * 1e6)));
    cudaFree(sink);
}

// ------------------------------------------------------------ FP64 butterfly
__device__ __forceinline__ double mulmod_f(double x, double w, double wq, double q) {
    const double M = 6755399441055744.0;                 // 1.5 * 2^52: (v + M) - M = rint(v)
    const double c = __dadd_rn(__fma_rn(x, wq, M), -M);
    const double h = __dmul_rn(x, w);
    const double l = __fma_rn(x, w, -h);
    const double r = __fma_rn(-c, q, h);
    return __dadd_rn(r, l);
}

template <int NV, int MINB, int VAR>
__global__ void __launch_bounds__(256, MINB) kb(unsigned iters, double q, double w0, double* sink, int flag) {
    double v[NV];
#pragma unroll
    for (int i = 0; i < NV; i++) v[i] = (double)((threadIdx.x * 1315423911ull + i * 2654435761ull + blockIdx.x) % 17592169062401ull);
    double w = w0 + threadIdx.x, wq = w / q;
    const double invq = 1.0 / q, M = 6755399441055744.0;
    for (unsigned it = 0; it < iters; it++) {
#pragma unroll
        for (int half = NV / 2; half >= 1; half >>= 1) {
#pragma unroll
            for (int j = 0; j < NV; j++) {
                if ((j & half) == 0) {
                    const int jj = j | half;
                    if (VAR == 0) {          // forward: X + wY, X - wY
                        const double T = mulmod_f(v[jj], w, wq, q);
                        const double X = v[j];
                        v[j] = __dadd_rn(X, T);
                        v[jj] = __dadd_rn(X, -T);
                    } else {                 // inverse: X + Y, (X - Y) w
                        const double X = v[j], Y = v[jj];
                        v[j] = __dadd_rn(X, Y);
                        v[jj] = mulmod_f(__dadd_rn(X, -Y), w, wq, q);
                    }
                }
            }
        }
        if (VAR == 1) {   // the inverse doubles the sum branch: reduce once per pass like the real kernel would
#pragma unroll
            for (int j = 0; j < NV; j++) {
                const double c = __dadd_rn(__fma_rn(v[j], invq, M), -M);
                v[j] = __fma_rn(-c, q, v[j]);
            }
        }
        w += 2.0; wq = w * invq;
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NV; i++) s += v[i];
    if (flag) sink[threadIdx.x] = s;
}

template <int NV, int MINB, int VAR>
void runb(const char* name, int sms, double mhz) {
    double* sink; cudaMalloc(&sink, 8 * 256);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const unsigned blocks = sms * MINB, iters = 3000;
    double best = 0;
    int stages = 0; for (int h = NV / 2; h >= 1; h >>= 1) stages++;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        kb<NV, MINB, VAR><<<blocks, 256>>>(iters, 17592169062401.0, 1299579534.0, sink, 0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double rate = (double)blocks * 256 * iters * (NV / 2) * stages / (ms * 1e-3);
        if (rep && rate > best) best = rate;
    }
    printf("%-34s %8.1f Gbutterfly/s -> %6.1f M NTT(n=4096)/s compute bound, %5.2f clk/SMSP per warp-butterfly\n", name,
           best / 1e9, best / 24576 / 1e6, sms * 4 * mhz * 1e6 * 32 / best);
    cudaFree(sink);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); const int s = p.multiProcessorCount;
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1000.0;
    printf("%s, %d SMs, %.0f MHz\n", p.name, s, mhz);
    run<0>("DFMA", s, mhz);
    run<1>("DADD", s, mhz);
    run<2>("DMUL", s, mhz);
    run<3>("DFMA + IMAD.lo", s, mhz);
    run<4>("DFMA + LOP3", s, mhz);
    run<5>("DFMA + IMAD.lo + LOP3", s, mhz);
    run<6>("IMAD.lo + LOP3", s, mhz);
    run<7>("2 DFMA + IMAD.WIDE", s, mhz);
    runb<16, 3, 0>("fp64 fwd butterfly NV=16 3cta", s, mhz);
    runb<16, 2, 0>("fp64 fwd butterfly NV=16 2cta", s, mhz);
    runb<16, 4, 0>("fp64 fwd butterfly NV=16 4cta", s, mhz);
    runb<8, 4, 0>("fp64 fwd butterfly NV=8 4cta", s, mhz);
    runb<16, 3, 1>("fp64 inv butterfly NV=16 3cta", s, mhz);
    return 0;
}
