// pointwise_variants.cu -- K3 (r = a * b mod q, 24 B per coefficient) formulations timed at the bench's working set
// (3 x 512 MiB, result aliasing b as in bench.py) to see which one gets closest to the copy peak.
// build: nvcc -O3 -std=c++17 -lineinfo -I lambda_snark_r_b200/csrc -I include -gencode arch=compute_100a,code=sm_100a \
//        tools/pointwise_variants.cu lambda_snark_r_b200/csrc/lsr_host.cpp -o tools/_bin/pw_variants
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "lsr_host.h"
#include "lsr_arith.cuh"
using namespace lsr;

// V0: the shipped form (one 16-byte pair per thread per iteration, default cache policy)
__global__ void __launch_bounds__(256) v0(const ModParams mp, u64* r, const u64* a, const u64* b, size_t total) {
    const size_t stride = (size_t)gridDim.x * blockDim.x, pairs = total >> 1;
    const ulonglong2* a2 = (const ulonglong2*)a; const ulonglong2* b2 = (const ulonglong2*)b; ulonglong2* r2 = (ulonglong2*)r;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < pairs; p += stride) {
        const ulonglong2 x = a2[p], y = b2[p];
        r2[p] = make_ulonglong2(mulmod_exact(x.x, y.x, mp), mulmod_exact(x.y, y.y, mp));
    }
}

// V1: U pairs per thread per iteration, all 2U loads issued before the first product; streaming hints optional
template <int U, bool CS>
__global__ void __launch_bounds__(256) v1(const ModParams mp, u64* r, const u64* a, const u64* b, size_t total) {
    const size_t pairs = total >> 1;
    const ulonglong2* a2 = (const ulonglong2*)a; const ulonglong2* b2 = (const ulonglong2*)b; ulonglong2* r2 = (ulonglong2*)r;
    const size_t tile = (size_t)blockDim.x * U;
    for (size_t base = (size_t)blockIdx.x * tile; base < pairs; base += (size_t)gridDim.x * tile) {
        ulonglong2 x[U], y[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const size_t p = base + (size_t)u * blockDim.x + threadIdx.x;
            if (p < pairs) { x[u] = CS ? __ldcs(a2 + p) : a2[p]; y[u] = CS ? __ldcs(b2 + p) : b2[p]; }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const size_t p = base + (size_t)u * blockDim.x + threadIdx.x;
            if (p < pairs) {
                const ulonglong2 v = make_ulonglong2(mulmod_exact(x[u].x, y[u].x, mp), mulmod_exact(x[u].y, y[u].y, mp));
                if (CS) __stcs(r2 + p, v); else r2[p] = v;
            }
        }
    }
}

// V2: as V1 but for q < 2^61 with the Barrett constants as immediates of the instantiation's kernel parameters is what
// mulmod_exact already does; this variant drops the Goldilocks branch test (mp.gold) from the inner loop
template <int U>
__global__ void __launch_bounds__(256) v2(const ModParams mp, u64* r, const u64* a, const u64* b, size_t total) {
    const size_t pairs = total >> 1;
    const ulonglong2* a2 = (const ulonglong2*)a; const ulonglong2* b2 = (const ulonglong2*)b; ulonglong2* r2 = (ulonglong2*)r;
    const size_t tile = (size_t)blockDim.x * U;
    for (size_t base = (size_t)blockIdx.x * tile; base < pairs; base += (size_t)gridDim.x * tile) {
        ulonglong2 x[U], y[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const size_t p = base + (size_t)u * blockDim.x + threadIdx.x;
            if (p < pairs) { x[u] = __ldcs(a2 + p); y[u] = __ldcs(b2 + p); }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const size_t p = base + (size_t)u * blockDim.x + threadIdx.x;
            if (p < pairs)
                __stcs(r2 + p, make_ulonglong2(barrett128(x[u].x * y[u].x, __umul64hi(x[u].x, y[u].x), mp),
                                               barrett128(x[u].y * y[u].y, __umul64hi(x[u].y, y[u].y), mp)));
        }
    }
}

__global__ void copy_k(ulonglong2* d, const ulonglong2* s, size_t pairs) {
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < pairs; p += (size_t)gridDim.x * blockDim.x) d[p] = s[p];
}

int main(int argc, char** argv) {
    const u64 q = 17592169062401ull;
    const size_t total = (size_t)(argc > 1 ? atoll(argv[1]) : 16384) * 4096;
    ModParams mp = host::make_mod_params(q, 12);
    u64 *a, *b; cudaMalloc(&a, total * 8); cudaMalloc(&b, total * 8);
    std::vector<u64> h(total);
    for (size_t i = 0; i < total; i++) h[i] = (i * 2654435761ull + 977) % q;
    cudaMemcpy(a, h.data(), total * 8, cudaMemcpyHostToDevice); cudaMemcpy(b, h.data(), total * 8, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    auto run = [&](const char* name, auto launch, double bytes_per_coeff) {
        float best = 1e9;
        for (int rep = 0; rep < 12; rep++) {
            cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 2 && ms < best) best = ms;
        }
        cudaError_t err = cudaGetLastError();
        printf("%-34s %.4f ms  %7.1f GB/s%s\n", name, best, total * bytes_per_coeff / best / 1e6, err ? cudaGetErrorString(err) : "");
    };
    run("copy a->b (16 B/coeff)", [&] { copy_k<<<sms * 16, 256>>>((ulonglong2*)b, (const ulonglong2*)a, total / 2); }, 16);
    cudaMemcpy(b, h.data(), total * 8, cudaMemcpyHostToDevice);
    for (int g : {8, 16, 32}) {
        char nm[64];
        snprintf(nm, 64, "v0 grid %dx SMs", g); run(nm, [&] { v0<<<sms * g, 256>>>(mp, b, a, b, total); }, 24);
        snprintf(nm, 64, "v1 U=2 grid %dx", g); run(nm, [&] { v1<2, false><<<sms * g, 256>>>(mp, b, a, b, total); }, 24);
        snprintf(nm, 64, "v1 U=4 grid %dx", g); run(nm, [&] { v1<4, false><<<sms * g, 256>>>(mp, b, a, b, total); }, 24);
        snprintf(nm, 64, "v1 U=4 cs grid %dx", g); run(nm, [&] { v1<4, true><<<sms * g, 256>>>(mp, b, a, b, total); }, 24);
        snprintf(nm, 64, "v1 U=8 cs grid %dx", g); run(nm, [&] { v1<8, true><<<sms * g, 256>>>(mp, b, a, b, total); }, 24);
        snprintf(nm, 64, "v2 U=4 grid %dx", g); run(nm, [&] { v2<4><<<sms * g, 256>>>(mp, b, a, b, total); }, 24);
    }
    // one CTA per tile, no grid-stride loop
    {
        const size_t pairs = total / 2;
        run("v1 U=4 cs one tile per CTA", [&] { v1<4, true><<<(unsigned)((pairs + 1023) / 1024), 256>>>(mp, b, a, b, total); }, 24);
        run("v1 U=8 cs one tile per CTA", [&] { v1<8, true><<<(unsigned)((pairs + 2047) / 2048), 256>>>(mp, b, a, b, total); }, 24);
    }
    return 0;
}
