// Micro-benchmark: sustained per-SM rates of the instructions the NTT butterfly can be built from
// (integer multiply forms vs FP64 FMA), to decide between the integer and the FP64 butterfly.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;

template <int MODE>
__global__ void k(u64 *out, u64 seed, int iters) {
    u64 a[8];
    double d[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = seed + threadIdx.x * 977 + i * 31; d[i] = (double)(a[i] & 0xffffff) + 0.5; }
    const unsigned m0 = (unsigned)seed | 1, m1 = (unsigned)(seed >> 32) | 1;
    const double c = 1.0000001, e = 0.9999999;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) { unsigned lo = (unsigned)a[i]; lo = lo * m0 + m1; a[i] = (a[i] & 0xffffffff00000000ull) | lo; }           // IMAD (32-bit lo)
            if (MODE == 1) { a[i] = (u64)(unsigned)a[i] * m0 + a[i]; }                                                                  // IMAD.WIDE
            if (MODE == 2) { unsigned lo = (unsigned)a[i]; lo = __umulhi(lo, m0) + m1; a[i] = (a[i] & 0xffffffff00000000ull) | lo; }    // IMAD.HI
            if (MODE == 3) { d[i] = __fma_rn(d[i], c, e); }                                                                             // DFMA
            if (MODE == 4) { d[i] = __dadd_rn(d[i], c); }                                                                               // DADD
            if (MODE == 5) { a[i] = a[i] + m0 + (a[i] >> 7); }                                                                          // IADD3/shift mix (alu)
        }
    }
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i] + (u64)d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, u64 *out) {
    const int iters = 4096, blocks = 148 * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, 12345, 16);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, 12345, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * threads * 8.0 * iters;
    printf("%-10s %8.3f ms  %7.1f Gop/s  %6.1f lanes/clk/SM (at 1.965 GHz)\n", name, ms, ops / ms / 1e6, ops / (ms * 1e-3) / 148 / 1.965e9);
}
int main() {
    u64 *out; cudaMalloc(&out, 148 * 8 * 256 * 8);
    run<0>("IMAD.lo", out); run<1>("IMAD.WIDE", out); run<2>("IMAD.HI", out); run<3>("DFMA", out); run<4>("DADD", out); run<5>("ALU mix", out);
    return 0;
}
