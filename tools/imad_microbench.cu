// imad_microbench.cu -- which IMAD forms run at 64 lanes/clk/SM on B200?
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/imad_microbench.cu -o gpurun_out/imad_mb
#include <cstdio>
#include <cuda_runtime.h>

constexpr int CH = 8;
constexpr int INNER = 32;

template <int OP>
__global__ void __launch_bounds__(256) k(unsigned iters, unsigned seed, unsigned long long* sink) {
    unsigned a[CH], b[CH], lo[CH];
    unsigned long long acc[CH];
#pragma unroll
    for (int c = 0; c < CH; c++) {
        a[c] = threadIdx.x * 2654435761u + seed + c * 77u;
        b[c] = blockIdx.x * 40503u + 1u + c * 1315423911u;
        lo[c] = a[c] ^ b[c];
        acc[c] = ((unsigned long long)a[c] << 32) | b[c];
    }
    for (unsigned it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < INNER; r++) {
#pragma unroll
            for (int c = 0; c < CH; c++) {
                if (OP == 0) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo[c]) : "r"(a[c]), "r"(b[c]));
                if (OP == 1) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[c]) : "r"(a[c]), "r"(b[c]));
                if (OP == 2) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(lo[c]) : "r"(a[c]), "r"(b[c]));
                if (OP == 3) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(lo[c]) : "r"(a[c]), "r"(b[c]));
                if (OP == 4) { unsigned t = (unsigned)acc[c]; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(acc[c]) : "r"(t), "r"(b[c])); }
                if (OP == 8) { unsigned t = (unsigned)acc[c]; asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[c]) : "r"(t), "r"(b[c])); }
                if (OP == 9) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(lo[c]) : "r"(a[c]));
                if (OP == 5) {   // wide product whose high half feeds the next op (mulhi via wide)
                    unsigned long long t;
                    asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(lo[c]), "r"(b[c]));
                    lo[c] = (unsigned)(t >> 32) + a[c];
                }
                if (OP == 6) asm volatile("add.u32 %0, %0, %1;" : "+r"(lo[c]) : "r"(a[c]));                 // IADD3 reference
                if (OP == 10) {  // one IMAD + one IADD3, independent chains
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(lo[c]) : "r"(a[c]), "r"(b[c]));
                    asm volatile("xor.b32 %0, %0, %1;" : "+r"(a[c]) : "r"(b[c]));
                }
                if (OP == 11) {  // one IMAD.WIDE + two IADD3-class (carry pair), like a butterfly
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[c]) : "r"(a[c]), "r"(b[c]));
                    asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(lo[c]), "+r"(a[c]) : "r"(b[c]), "r"(b[(c + 1) % CH]));
                }
                if (OP == 12) asm volatile("xor.b32 %0, %0, %1;" : "+r"(lo[c]) : "r"(a[c]));                // LOP3 reference
                if (OP == 7) asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;"                      // 64-bit add
                                          : "+r"(lo[c]), "+r"(a[c]) : "r"(b[c]), "r"(b[(c + 1) % CH]));
            }
        }
    }
    unsigned long long s = 0;
#pragma unroll
    for (int c = 0; c < CH; c++) s += acc[c] + lo[c] + a[c];
    if (s == 0x123456789abcdefULL) *sink = s;
}

template <int OP>
double run(const char* name, int sms, double ops_per_inst) {
    unsigned long long* sink;
    cudaMalloc(&sink, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const unsigned blocks = sms * 8, iters = 4000;
    double best = 0;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        k<OP><<<blocks, 256>>>(iters, rep, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double rate = (double)blocks * 256 * iters * INNER * CH * ops_per_inst / (ms * 1e-3) / 1e9;
        if (rep && rate > best) best = rate;
    }
    printf("%-28s %10.0f Gop/s  = %6.1f lanes/clk/SM @1.965GHz\n", name, best, best * 1e9 / (sms * 1.965e9));
    cudaFree(sink);
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    int s = p.multiProcessorCount;
    run<0>("mad.lo  d=a*b+d", s, 1);
    run<1>("mad.wide D=a*b+D", s, 1);
    run<2>("mad.hi  d=hi(d*a)+b", s, 1);
    run<3>("mad.lo  d=d*a+b", s, 1);
    run<4>("mul.wide D=lo(D)*b", s, 1);
    run<8>("mad.wide D=lo(D)*b+D", s, 1);
    run<9>("mul.hi d=hi(d*a)", s, 1);
    run<5>("mul.wide + hi + add", s, 1);
    run<6>("add.u32", s, 1);
    run<7>("add.cc/addc (2 inst)", s, 2);
    run<12>("xor.b32", s, 1);
    run<10>("mad.lo + xor (2 inst)", s, 2);
    run<11>("mad.wide + add.cc/addc (3 inst)", s, 3);
    return 0;
}
