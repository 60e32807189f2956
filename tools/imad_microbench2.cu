// imad_microbench2.cu -- IMAD-form throughput with operands that change every instruction
// (nothing for ptxas to hoist or strength-reduce).  Prints SASS-verified rates.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int CH = 8, INNER = 32;
template <int OP>
__global__ void __launch_bounds__(256) k(unsigned iters, unsigned seed, unsigned long long* sink, int flag) {
    unsigned b[CH], d[CH];
    unsigned long long D[CH];
#pragma unroll
    for (int c = 0; c < CH; c++) {
        b[c] = (blockIdx.x * 40503u + threadIdx.x * 2654435761u + c * 1315423911u + seed) | 1u;
        d[c] = b[c] * 7u + c;
        D[c] = ((unsigned long long)d[c] << 32) | b[c];
    }
    for (unsigned it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < INNER; r++) {
#pragma unroll
            for (int c = 0; c < CH; c++) {
                const int n = (c + 1) % CH;
                if (OP == 0) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(D[c]) : "r"((unsigned)D[n]), "r"(b[c]));
                if (OP == 1) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(D[c]) : "r"((unsigned)D[n]), "r"(b[c]));
                if (OP == 2) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(d[c]) : "r"(d[n]), "r"(b[c]));
                if (OP == 3) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(d[c]) : "r"(d[n]), "r"(b[c]));
                if (OP == 4) asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(d[c]) : "r"(d[n]), "r"(b[c]));
                if (OP == 5) asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(d[c]), "+r"(b[c]) : "r"(d[n]), "r"(b[n]));
                if (OP == 6) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(d[c]) : "r"(d[n]), "r"(b[c]));
                if (OP == 7) {  // IMAD.WIDE(acc) + independent LOP3: do the pipes overlap?
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(D[c]) : "r"((unsigned)D[n]), "r"(b[c]));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(d[c]) : "r"(d[n]), "r"(b[c]));
                }
                if (OP == 8) {  // IMAD.lo + independent LOP3
                    asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(d[c]) : "r"(d[n]), "r"(b[c]));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[c]) : "r"(b[n]), "r"((unsigned)D[c]));
                }
            }
        }
    }
    unsigned long long s = 0;
#pragma unroll
    for (int c = 0; c < CH; c++) s += D[c] + d[c] + b[c];
    if (flag) sink[threadIdx.x] = s;
}
template <int OP> void run(const char* name, int sms, double inst_per_op) {
    unsigned long long* sink; cudaMalloc(&sink, 8 * 256);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const unsigned blocks = sms * 8, iters = 3000; double best = 0;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0); k<OP><<<blocks, 256>>>(iters, rep, sink, 0); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double rate = (double)blocks * 256 * iters * INNER * CH / (ms * 1e-3);
        if (rep && rate > best) best = rate;
    }
    printf("%-44s %7.2f ops/clk/SM  (%4.2f clk per warp-op per SMSP; %g inst/op)\n", name, best / (sms * 1.965e9),
           4 * 32 / (best / (sms * 1.965e9)), inst_per_op);
    cudaFree(sink);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int s = p.multiProcessorCount;
    run<0>("IMAD.WIDE  D = a*b + D (64-bit addend)", s, 1);
    run<1>("IMAD.WIDE  D = a*b", s, 1);
    run<2>("IMAD       d = a*b + d", s, 1);
    run<3>("IMAD.HI    d = hi(a*b) + d", s, 1);
    run<4>("IMAD       d = a*b", s, 1);
    run<5>("IADD3 + IADD3.X (64-bit add)", s, 2);
    run<6>("LOP3       d = d^a^b", s, 1);
    run<7>("IMAD.WIDE(acc) + LOP3", s, 2);
    run<8>("IMAD + LOP3", s, 2);
    return 0;
}
