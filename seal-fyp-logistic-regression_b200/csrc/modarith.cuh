// modarith.cuh -- 64-bit modular arithmetic for RNS limbs on sm_100a.
//
// Everything is unsigned 64-bit integer work built from 32-bit IMADs (__umul64hi); there is
// no floating point and no tensor-core path (modular arithmetic is not a dense contraction).
// Primes are < 2^62 so the lazy ranges [0,2p) / [0,4p) fit in a word.
#pragma once
#include <cstdint>

typedef unsigned long long u64;

struct ModConst {
    u64 p;      // prime (at most 60 bits, as SEAL requires of user primes)
    u64 p2;     // 2p
    u64 r0;     // floor(2^128 / p), low word
    u64 r1;     // floor(2^128 / p), high word  (== floor(2^64 / p))
    u64 ninv;   // N^-1 mod p
    u64 ninvs;  // Shoup companion of ninv
    u64 w1ni;   // (inverse twiddle of the root node) * N^-1 mod p
    u64 w1nis;  // its Shoup companion
    u64 gsc;    // smallest multiple of p that is >= 4p + 2^47 (offset of the lazy inverse butterfly)
    u64 p4;     // 4p
    u64 negp;   // 2^64 - p (kept opaque so that w*x + q*negp stays one multiply-add chain)
    u64 p4hi;   // high 32 bits of 4p (threshold of the cheap lazy correction)
};

// w*x mod p, lazily reduced to [0,4p); ws = floor(w * 2^64 / p), any 64-bit x, w < p.
// Written as one PTX sequence of 9 integer multiply-adds (the multiplier pipe is the binding
// resource):
//   q = truncated high product of ws*x: a1*b1 + hi32(a1*b0) + hi32(a0*b1) -- at most 2 below the
//       true floor(ws*x / 2^64); the Shoup remainder absorbs the error ([0,4p) instead of [0,2p))
//   r = low 64 bits of w*x + q*negp,  negp = 2^64 - p
__device__ __forceinline__ u64 shoup_lazy(u64 x, u64 w, u64 ws, u64 negp) {
    const unsigned x0 = (unsigned)x, x1 = (unsigned)(x >> 32);
    const unsigned w0 = (unsigned)w, w1 = (unsigned)(w >> 32);
    const unsigned s0 = (unsigned)ws, s1 = (unsigned)(ws >> 32);
    const unsigned n0 = (unsigned)negp, n1 = (unsigned)(negp >> 32);
    unsigned rlo, rhi;
    asm("{\n\t"
        ".reg .u32 q0, q1, tl, th;\n\t"
        ".reg .u64 t;\n\t"
        "mul.wide.u32    t, %3, %7;\n\t"          // s1*x1
        "mov.b64         {q0, q1}, t;\n\t"
        "mad.hi.cc.u32   q0, %3, %6, q0;\n\t"     // + hi32(s1*x0)
        "addc.u32        q1, q1, 0;\n\t"
        "mad.hi.cc.u32   q0, %2, %7, q0;\n\t"     // + hi32(s0*x1)
        "addc.u32        q1, q1, 0;\n\t"
        "mul.wide.u32    t, %4, %6;\n\t"          // w0*x0
        "mad.wide.u32    t, q0, %8, t;\n\t"       // + q0*n0
        "mov.b64         {tl, th}, t;\n\t"
        "mad.lo.u32      th, %4, %7, th;\n\t"     // + (w0*x1) << 32
        "mad.lo.u32      th, %5, %6, th;\n\t"     // + (w1*x0) << 32
        "mad.lo.u32      th, q0, %9, th;\n\t"     // + (q0*n1) << 32
        "mad.lo.u32      th, q1, %8, th;\n\t"     // + (q1*n0) << 32
        "mov.u32         %0, tl;\n\t"
        "mov.u32         %1, th;\n\t"
        "}"
        : "=r"(rlo), "=r"(rhi)
        : "r"(s0), "r"(s1), "r"(w0), "r"(w1), "r"(x0), "r"(x1), "r"(n0), "r"(n1));
    return ((u64)rhi << 32) | rlo;
}
// exact variant, [0,2p)
__device__ __forceinline__ u64 shoup_lazy2(u64 x, u64 w, u64 ws, u64 p) {
    u64 q = __umul64hi(ws, x);
    return w * x - q * p;
}
// lazy correction with one 32-bit compare and a predicated subtract: removes 4p only when the high
// word proves x > 4p.   [0, 8p + 2^32) -> [0, 4p + 2^32); exactness is restored by the final
// Barrett reduction of each transform.
__device__ __forceinline__ u64 lazy_sub(u64 x, u64 p4, unsigned p4hi) {
    unsigned lo = (unsigned)x, hi = (unsigned)(x >> 32);
    asm("{\n\t.reg .pred q;\n\tsetp.gt.u32 q, %1, %2;\n\t@q sub.cc.u32 %0, %0, %3;\n\t@q subc.u32 %1, %1, %4;\n\t}"
        : "+r"(lo), "+r"(hi)
        : "r"(p4hi), "r"((unsigned)p4), "r"((unsigned)(p4 >> 32)));
    return ((u64)hi << 32) | lo;
}
__device__ __forceinline__ u64 csub(u64 x, u64 p) { return x >= p ? x - p : x; }
__device__ __forceinline__ u64 shoup_mul(u64 x, u64 w, u64 ws, u64 p) { return csub(shoup_lazy2(x, w, ws, p), p); }
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

// 128-bit accumulate: (lo,hi) += a*b.  The four 32x32 partial products are added straight into
// the accumulator words with one carry chain (4 wide multiplies, no separate lo/hi products).
__device__ __forceinline__ void mac128(u64 &lo, u64 &hi, u64 a, u64 b) {
    unsigned a0 = (unsigned)a, a1 = (unsigned)(a >> 32), b0 = (unsigned)b, b1 = (unsigned)(b >> 32);
    unsigned r0 = (unsigned)lo, r1 = (unsigned)(lo >> 32), r2 = (unsigned)hi, r3 = (unsigned)(hi >> 32);
    asm("{\n\t"
        "mad.lo.cc.u32   %0, %4, %6, %0;\n\t"     // a0*b0 lo -> r0
        "madc.hi.cc.u32  %1, %4, %6, %1;\n\t"     // a0*b0 hi -> r1
        "madc.lo.cc.u32  %2, %5, %7, %2;\n\t"     // a1*b1 lo -> r2
        "madc.hi.u32     %3, %5, %7, %3;\n\t"     // a1*b1 hi -> r3
        "mad.lo.cc.u32   %1, %4, %7, %1;\n\t"     // a0*b1 lo -> r1
        "madc.hi.cc.u32  %2, %4, %7, %2;\n\t"     // a0*b1 hi -> r2
        "addc.u32        %3, %3, 0;\n\t"
        "mad.lo.cc.u32   %1, %5, %6, %1;\n\t"     // a1*b0 lo -> r1
        "madc.hi.cc.u32  %2, %5, %6, %2;\n\t"     // a1*b0 hi -> r2
        "addc.u32        %3, %3, 0;\n\t"
        "}"
        : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3)
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    lo = ((u64)r1 << 32) | r0;
    hi = ((u64)r3 << 32) | r2;
}

// Harvey-style butterflies with relaxed ranges -----------------------------------------------------
// forward (Cooley-Tukey): X,Y in [0, 8p + 2^32) -> [0, 8p + 2^32)
__device__ __forceinline__ void ct_bfly(u64 &X, u64 &Y, u64 w, u64 ws, const ModConst &m) {
    u64 x = lazy_sub(X, m.p4, (unsigned)m.p4hi);   // < 4p + 2^32
    u64 t = shoup_lazy(Y, w, ws, m.negp);          // < 4p
    X = x + t;
    Y = x + m.p4 - t;
}
// inverse (Gentleman-Sande): after s stages X < 4p + 2^(31+s), Y < 4p; gsc (a multiple of p, at
// least 4p + 2^47) keeps the difference non-negative
__device__ __forceinline__ void gs_bfly(u64 &X, u64 &Y, u64 w, u64 ws, const ModConst &m) {
    u64 s = X + Y;
    u64 d = X + m.gsc - Y;
    X = lazy_sub(s, m.p4, (unsigned)m.p4hi);
    Y = shoup_lazy(d, w, ws, m.negp);
}

// ================================================================================================
// FP64 butterflies for primes below 2^41.
// On B200 a warp-wide IMAD.WIDE occupies the integer multiplier for 8 cycles and IMAD.HI for 4
// (16 resp. 32 lanes/clk/SM, profiles/micro/pipe_rates.cu), while DFMA / DADD / DMUL run at 64
// lanes/clk/SM on their own pipe.  For small primes the butterfly is therefore done on integer-valued
// doubles: a modular product costs 7 FP64 operations, all exact:
//     h = RN(w*y),  l = w*y - h (FMA, exact error term),  q = rint(h/p),  r = h - q*p (FMA, exact
//     because |h - q p| < 1.01 p is representable),  t = r + l,   |t| < 2p.
// Values are signed and lazily bounded (|x| < 2^51 keeps every quantity an exact integer): forward
// butterflies add at most 2p per stage, inverse sums double per stage and are reduced once per pass.
// Results are bit-identical to the integer path because only exact integer arithmetic is used.
struct FpConst {
    double p, pinv, ninv, w1ni, ok, c31;   // c31 = 2^31 mod p
};
__device__ __forceinline__ double fp_rint(double v) {   // round to nearest integer, |v| < 2^51
    const double M = 6755399441055744.0;                 // 1.5 * 2^52
    return __dadd_rn(__dadd_rn(v, M), -M);
}
// rint(a*b) with the product, the magic-number add and the rounding in one FMA, |a*b| < 2^51
__device__ __forceinline__ double fp_rint_mul(double a, double b) {
    const double M = 6755399441055744.0;
    return __dadd_rn(__fma_rn(a, b, M), -M);
}
__device__ __forceinline__ double fp_mulmod(double y, double w, const FpConst &f) {
    double h = __dmul_rn(w, y);
    double l = __fma_rn(w, y, -h);
    double q = fp_rint_mul(h, f.pinv);
    double r = __fma_rn(-q, f.p, h);
    return __dadd_rn(r, l);
}
// x -> representative in about (-p/2, p/2)
__device__ __forceinline__ double fp_reduce(double x, const FpConst &f) {
    double q = fp_rint_mul(x, f.pinv);
    return __fma_rn(-q, f.p, x);
}
// exact conversions for integers below 2^52
__device__ __forceinline__ double fp_from_u64(u64 v) {
    return __dadd_rn(__longlong_as_double((long long)(v | 0x4330000000000000ull)), -4503599627370496.0);
}
// canonical residue of a large prime (v < 2^62) -> a double congruent to it modulo the small prime, |.| < 2p + 2^31
__device__ __forceinline__ double fp_reduce_big(u64 v, const FpConst &f) {
    const double hi = fp_from_u64(v >> 31), lo = fp_from_u64(v & 0x7fffffffull);
    return __dadd_rn(fp_mulmod(hi, f.c31, f), lo);
}
__device__ __forceinline__ u64 fp_to_canonical(double x, const FpConst &f) {   // -> [0,p)
    double v = fp_reduce(x, f);
    if (v < 0.0) v = __dadd_rn(v, f.p);
    return (u64)__double_as_longlong(__dadd_rn(v, 4503599627370496.0)) & 0x000fffffffffffffull;
}
__device__ __forceinline__ void ct_bfly_fp(double &X, double &Y, double w, const FpConst &f) {
    double t = fp_mulmod(Y, w, f);
    Y = __dadd_rn(X, -t);
    X = __dadd_rn(X, t);
}
__device__ __forceinline__ void gs_bfly_fp(double &X, double &Y, double w, const FpConst &f) {
    double s = __dadd_rn(X, Y);
    double d = __dadd_rn(X, -Y);
    X = s;
    Y = fp_mulmod(d, w, f);
}
