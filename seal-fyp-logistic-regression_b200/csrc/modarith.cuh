// modarith.cuh -- 64-bit modular arithmetic for RNS limbs on sm_100a.
//
// Everything is unsigned 64-bit integer work built from 32-bit IMADs (__umul64hi); there is
// no floating point and no tensor-core path (modular arithmetic is not a dense contraction).
// Primes are < 2^62 so the lazy ranges [0,2p) / [0,4p) fit in a word.
#pragma once
#include <cstdint>

typedef unsigned long long u64;

struct ModConst {
    u64 p;      // prime
    u64 p2;     // 2p
    u64 r0;     // floor(2^128 / p), low word
    u64 r1;     // floor(2^128 / p), high word  (== floor(2^64 / p))
    u64 ninv;   // N^-1 mod p
    u64 ninvs;  // Shoup companion of ninv
    u64 w1ni;   // (inverse twiddle of the root node) * N^-1 mod p
    u64 w1nis;  // its Shoup companion
};

// w*x mod p, lazily reduced to [0,2p); ws = floor(w * 2^64 / p), any 64-bit x, w < p
__device__ __forceinline__ u64 shoup_lazy(u64 x, u64 w, u64 ws, u64 p) {
    u64 q = __umul64hi(ws, x);
    return w * x - q * p;
}
__device__ __forceinline__ u64 csub(u64 x, u64 p) { return x >= p ? x - p : x; }
__device__ __forceinline__ u64 shoup_mul(u64 x, u64 w, u64 ws, u64 p) { return csub(shoup_lazy(x, w, ws, p), p); }
__device__ __forceinline__ u64 addmod(u64 a, u64 b, u64 p) { return csub(a + b, p); }
__device__ __forceinline__ u64 submod(u64 a, u64 b, u64 p) { return a >= b ? a - b : a + p - b; }

// any 64-bit x -> [0,p)
__device__ __forceinline__ u64 reduce64(u64 x, const ModConst &m) {
    u64 q = __umul64hi(x, m.r1);
    return csub(x - q * m.p, m.p);
}

// (lo,hi) < 2^128 -> [0,p); two-word Barrett with ratio floor(2^128/p)
__device__ __forceinline__ u64 barrett128(u64 lo, u64 hi, const ModConst &m) {
    // q = floor((hi*2^64 + lo) * (r1*2^64 + r0) / 2^128), low word only (q < 2^64 since x/p < 2^64
    // is not required: we only need q mod 2^64 for the final subtraction)
    u64 carry = __umul64hi(lo, m.r0);
    u64 t_lo = lo * m.r1, t_hi = __umul64hi(lo, m.r1);
    u64 s = t_lo + carry;
    u64 tmp3 = t_hi + (s < t_lo);
    u64 u_lo = hi * m.r0, u_hi = __umul64hi(hi, m.r0);
    u64 s2 = s + u_lo;
    u64 c2 = u_hi + (s2 < s);
    u64 q = hi * m.r1 + tmp3 + c2;
    return csub(lo - q * m.p, m.p);
}
__device__ __forceinline__ u64 mulmod(u64 a, u64 b, const ModConst &m) {
    return barrett128(a * b, __umul64hi(a, b), m);
}

// 128-bit accumulate: (lo,hi) += a*b
__device__ __forceinline__ void mac128(u64 &lo, u64 &hi, u64 a, u64 b) {
    u64 pl = a * b, ph = __umul64hi(a, b);
    lo += pl;
    hi += ph + (lo < pl);
}

// Harvey butterflies --------------------------------------------------------------------------
// forward (Cooley-Tukey): X,Y in [0,4p) -> [0,4p)
__device__ __forceinline__ void ct_bfly(u64 &X, u64 &Y, u64 w, u64 ws, u64 p, u64 p2) {
    u64 x = X >= p2 ? X - p2 : X;
    u64 t = shoup_lazy(Y, w, ws, p);
    X = x + t;
    Y = x + p2 - t;
}
// inverse (Gentleman-Sande): X,Y in [0,2p) -> [0,2p)
__device__ __forceinline__ void gs_bfly(u64 &X, u64 &Y, u64 w, u64 ws, u64 p, u64 p2) {
    u64 s = X + Y;
    u64 d = X + p2 - Y;
    X = s >= p2 ? s - p2 : s;
    Y = shoup_lazy(d, w, ws, p);
}
