// Micro-benchmark: the empirical COMPUTE ceiling of the engine's own butterflies on one B200.
// The radix-8 register steps of ntt_passes.cuh (fwd8 / inv8 for 60-bit primes, fwd8d / inv8d for primes below 2^41) run on
// register-resident data with register-resident twiddles -- no global or shared memory traffic, no barriers, no index
// arithmetic -- at the occupancies the transform kernels have (4 CTAs of 256 threads per SM) and at full occupancy.
// butterflies/s from here x the butterfly count of a key switch = the time below which no scheduling of these
// instruction sequences can go; bench.py and DESIGN.md quote the result (profiles/r02_butterfly_rates.txt).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I seal-fyp-logistic-regression_b200/csrc \
//        -o profiles/micro/butterfly_rates profiles/micro/butterfly_rates.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "modarith.cuh"
#include "ntt_passes.cuh"

// experiment: truncated Shoup product whose two cross terms hi32(s1*x0) + hi32(s0*x1) come from the FP64 pipe
// (round-toward-zero products: never above the true value, at most 1 below), leaving 3 wide + 4 low multiplies
__device__ __forceinline__ double u32_to_double(unsigned v) { return __hiloint2double(0x43300000, (int)v) - 4503599627370496.0; }
__device__ __forceinline__ u64 shoup_lazy_v2(u64 x, u64 w, u64 ws, u64 negp) {
    const unsigned x0 = (unsigned)x, x1 = (unsigned)(x >> 32);
    const unsigned s0 = (unsigned)ws, s1 = (unsigned)(ws >> 32);
    const double mid = __fma_rz(u32_to_double(x0), u32_to_double(s1), __dmul_rz(u32_to_double(x1), u32_to_double(s0)));
    const u64 tb = (u64)__double_as_longlong(__fma_rz(mid, 0x1p-32, 4503599627370496.0));
    const u64 q = (u64)x1 * s1 + (tb & 0x3ffffffffull);
    return w * x + q * negp;
}
__device__ __forceinline__ void ct_bfly_v2(u64 &X, u64 &Y, u64 w, u64 ws, const ModConst &m) {
    u64 x = lazy_sub(X, m.p4, (unsigned)m.p4hi);
    u64 t = shoup_lazy_v2(Y, w, ws, m.negp);
    X = x + t;
    Y = x + m.p4 - t;
}
__device__ __forceinline__ void fwd8_v2(u64 (&x)[8], const Tw8 &t, const ModConst &m) {
#pragma unroll
    for (int e = 0; e < 4; e++) ct_bfly_v2(x[e], x[e + 4], t.w[0].x, t.w[0].y, m);
    ct_bfly_v2(x[0], x[2], t.w[1].x, t.w[1].y, m);
    ct_bfly_v2(x[1], x[3], t.w[1].x, t.w[1].y, m);
    ct_bfly_v2(x[4], x[6], t.w[2].x, t.w[2].y, m);
    ct_bfly_v2(x[5], x[7], t.w[2].x, t.w[2].y, m);
#pragma unroll
    for (int q = 0; q < 4; q++) ct_bfly_v2(x[2 * q], x[2 * q + 1], t.w[3 + q].x, t.w[3 + q].y, m);
}

template <int MODE, int OCC>
__global__ void __launch_bounds__(256, OCC) k_bfly(u64 *out, ModConst m, FpConst f, u64 seed, int iters) {
    u64 x[8];
    double xd[8];
    Tw8 t;
    Tw8d td;
#pragma unroll
    for (int e = 0; e < 8; e++) {
        x[e] = (seed * (threadIdx.x + 1) + e * 0x9E3779B97F4A7C15ull) % m.p;
        xd[e] = (double)((seed * (threadIdx.x + 3) + e * 977) % (u64)f.p);
    }
#pragma unroll
    for (int q = 0; q < 7; q++) {
        const u64 w = (seed * (q + 5) + threadIdx.x) % m.p;
        t.w[q].x = w;
        t.w[q].y = (u64)(((unsigned __int128)w << 64) / m.p);
        td.w[q] = (double)((seed * (q + 11) + threadIdx.x) % (u64)f.p);
    }
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) fwd8<0>(x, t, m);                     // 12 integer Cooley-Tukey butterflies
        if (MODE == 1) inv8<0>(x, t, m);                     // 12 integer Gentleman-Sande butterflies
        if (MODE == 2) fwd8d<0>(xd, td, f);                  // 12 FP64 butterflies (lazy: reduce as the passes do)
        if (MODE == 3) inv8d<0>(xd, td, f);
        if (MODE == 4) fwd8_v2(x, t, m);                     // experiment: cross terms of the Shoup quotient on the FP64 pipe
        if (MODE == 2 || MODE == 3) {
            if ((it & 3) == 3) {
#pragma unroll
                for (int e = 0; e < 8; e++) xd[e] = fp_reduce(xd[e], f);   // once per 12 stages (the passes: once per 6-9)
            }
        }
        if (MODE == 1 && (it & 3) == 3) {
#pragma unroll
            for (int e = 0; e < 8; e++) x[e] = reduce64(x[e], m);   // the inverse butterfly's lazy range lasts 16 stages
        }
    }
    u64 s = 0;
#pragma unroll
    for (int e = 0; e < 8; e++) s += x[e] + (u64)(long long)xd[e];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static ModConst make_mod(u64 p) {
    ModConst m{};
    m.p = p;
    m.p2 = 2 * p;
    unsigned __int128 r = (~(unsigned __int128)0) / p;
    m.r0 = (u64)r;
    m.r1 = (u64)(r >> 64);
    const u64 need = 4 * p + (1ull << 47);
    m.gsc = ((need + p - 1) / p) * p;
    m.p4 = 4 * p;
    m.negp = 0 - p;
    m.p4hi = (4 * p) >> 32;
    return m;
}

template <int MODE, int OCC>
static double run(const char *name, u64 *out, const ModConst &m, const FpConst &f) {
    const int iters = 2048, blocks = 148 * OCC * 4, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_bfly<MODE, OCC><<<blocks, threads>>>(out, m, f, 12345, 16);
    cudaEventRecord(e0);
    k_bfly<MODE, OCC><<<blocks, threads>>>(out, m, f, 12345, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double bf = (double)blocks * threads * 12.0 * iters;
    const double rate = bf / (ms * 1e-3);
    int regs = 0;
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, k_bfly<MODE, OCC>) == cudaSuccess) regs = fa.numRegs;
    printf("%-34s occupancy %d CTAs/SM, %3d regs: %8.1f G butterflies/s  (%.3f us per N=32768 transform of 245760 butterflies)\n", name, OCC,
           regs, rate / 1e9, 245760.0 / rate * 1e6);
    return rate;
}

int main() {
    u64 *out;
    cudaMalloc(&out, (size_t)148 * 8 * 4 * 256 * 8);
    const ModConst m = make_mod(1152921504606584833ull);   // the 60-bit prime of {60, 40 x 8, 60} at N = 32768
    FpConst f{};
    f.p = 1099510054913.0;                                  // a 40-bit NTT prime
    f.pinv = 1.0 / f.p;
    f.ok = 1.0;
    double r[8];
    r[0] = run<0, 4>("integer forward (fwd8, 60-bit)", out, m, f);
    r[1] = run<0, 8>("integer forward (fwd8, 60-bit)", out, m, f);
    r[2] = run<1, 4>("integer inverse (inv8, 60-bit)", out, m, f);
    r[3] = run<1, 8>("integer inverse (inv8, 60-bit)", out, m, f);
    r[4] = run<2, 4>("FP64 forward (fwd8d, 40-bit)", out, m, f);
    r[5] = run<2, 8>("FP64 forward (fwd8d, 40-bit)", out, m, f);
    r[6] = run<3, 4>("FP64 inverse (inv8d, 40-bit)", out, m, f);
    r[7] = run<3, 8>("FP64 inverse (inv8d, 40-bit)", out, m, f);
    run<4, 4>("integer forward, FP64 cross terms", out, m, f);
    run<4, 8>("integer forward, FP64 cross terms", out, m, f);
    // one Galois key switch at N = 32768, L = 3, chain {60, 40, 40, ..., 60}: 20 transforms of 245760 butterflies --
    // integer: 4 forward into P/q0 of the mod-up + ... see DESIGN.md section 4: 7 forward + 3 inverse integer, 8 forward + 2 inverse FP64
    const double bi_f = 7 * 245760.0, bi_i = 3 * 245760.0, bf_f = 8 * 245760.0, bf_i = 2 * 245760.0;
    for (int o = 0; o < 2; o++) {
        const double ti = bi_f / r[0 + o] + bi_i / r[2 + o], tf = bf_f / r[4 + o] + bf_i / r[6 + o];
        printf("key switch (N = 32768, L = 3), butterflies only, %d CTAs/SM: integer %.2f us + FP64 %.2f us = %.2f us if the two pipes never "
               "overlap, %.2f us if they overlap perfectly\n", o ? 8 : 4, ti * 1e6, tf * 1e6, (ti + tf) * 1e6, (ti > tf ? ti : tf) * 1e6);
    }
    // both kinds at once (two streams, equal butterfly counts -- the 10 : 10 mix of L = 3), CTAs of both kernels co-resident
    {
        // 2 + 2 CTAs per SM: one wave of each kernel, co-resident (a full grid of either kernel would fill the register file)
        const int iters = 8192, blocks = 148 * 2, threads = 256;
        cudaStream_t s0, s1;
        cudaStreamCreate(&s0);
        cudaStreamCreate(&s1);
        cudaEvent_t e0, e1, j1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventCreate(&j1);
        u64 *out2;
        cudaMalloc(&out2, (size_t)blocks * threads * 8);
        cudaDeviceSynchronize();
        cudaEventRecord(e0, s0);
        cudaStreamWaitEvent(s1, e0, 0);
        k_bfly<0, 4><<<blocks, threads, 0, s0>>>(out, m, f, 12345, iters);
        k_bfly<2, 4><<<blocks, threads, 0, s1>>>(out2, m, f, 12345, iters);
        cudaEventRecord(j1, s1);
        cudaStreamWaitEvent(s0, j1, 0);
        cudaEventRecord(e1, s0);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double bf = 2.0 * blocks * threads * 12.0 * iters;
        // the same two launches one after the other, for the comparison
        cudaEvent_t f0, f1;
        cudaEventCreate(&f0);
        cudaEventCreate(&f1);
        cudaDeviceSynchronize();
        cudaEventRecord(f0, s0);
        k_bfly<0, 4><<<blocks, threads, 0, s0>>>(out, m, f, 12345, iters);
        k_bfly<2, 4><<<blocks, threads, 0, s0>>>(out2, m, f, 12345, iters);
        cudaEventRecord(f1, s0);
        cudaEventSynchronize(f1);
        float ms_serial;
        cudaEventElapsedTime(&ms_serial, f0, f1);
        printf("2 + 2 CTAs per SM: concurrent %.3f ms, back to back %.3f ms\n", ms, ms_serial);
        printf("integer forward + FP64 forward concurrently (two streams): %.1f G butterflies/s -> %.2f us per key switch of 20 x 245760 "
               "butterflies\n", bf / (ms * 1e-3) / 1e9, 20 * 245760.0 / (bf / (ms * 1e-3)) * 1e6);
    }
    return 0;
}
