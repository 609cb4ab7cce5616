// ntt_passes.cuh -- negacyclic NTT building blocks for one RNS limb of N = 2^LOGN words.
//
// A limb is viewed as an N1 x N2 row-major matrix (N1 = 64 rows, N2 = N/64 columns):
//
//   forward (Cooley-Tukey, natural in -> bit-reversed out, SEAL's order)
//     column pass : stages 0..5   act along the rows index r for a fixed column c
//     row pass    : stages 6..n-1 act inside each contiguous row of N2 words
//   inverse (Gentleman-Sande) runs the same stages backwards: row pass, then column pass.
//
// Every CTA is 256 threads and owns 2048 words (16 KB of shared memory): a 64 x 32 column
// tile, or 2048/N2 whole rows.  Each thread keeps 8 words in registers and performs up to
// three radix-2 stages (a radix-8 step) between shared-memory exchanges, so a column pass
// has one exchange and a row pass two.  Twiddles follow the binary tree of SEAL's table:
// the butterfly of stage s on global index g uses node 2^s + (g >> (n - s)); a radix-8 step
// rooted at node nd touches nodes nd, 2nd..2nd+1, 4nd..4nd+3.
//
// The passes are device functions working on a register array so that callers can fuse their
// own loads (Galois gather, base conversion) and epilogues (key inner product, mod-down).
#pragma once
#include "modarith.cuh"

#define NTT_THREADS 256
#define NTT_TILE 2048

template <int LOGN>
struct NttGeo {
    static constexpr int N = 1 << LOGN;
    static constexpr int N2LOG = LOGN - 6;
    static constexpr int N2 = 1 << N2LOG;  // words per row
    static constexpr int T = N2 / 8;       // threads per row in the row pass
    static constexpr int REM = N2LOG - 6;  // stages left for the third radix step (0..3)
    static constexpr int ROW_TILES = N / NTT_TILE;
    static constexpr int COL_TILES = N2 / 32;
    static_assert(LOGN >= 12 && LOGN <= 15, "supported degrees: 4096..32768");
};

typedef ulonglong2 tw_t;  // {w, floor(w 2^64 / p)}

// thread index inside the 256-thread group that owns a tile (the tile kernels are 256-thread CTAs: identity there; the
// limb-per-CTA kernel runs four such groups in one 1024-thread CTA)
__device__ __forceinline__ unsigned ntt_tid() { return threadIdx.x & (NTT_THREADS - 1); }

// swizzled shared-memory index for the row pass: keeps every access pattern of the three radix
// steps conflict-free for 64-bit words (bank pair = low 4 bits)
__device__ __forceinline__ int sw(int li) { return li ^ ((li >> 3) & 15); }

// ---- radix-8 register steps ------------------------------------------------------------------
// element e of x sits at global index base + e*stride; SKIP leading (coarse) stages are omitted.
// Twiddle loading is split from the arithmetic so that callers can issue the loads of the next
// step before a shared-memory exchange and have them in flight across the barrier.
struct Tw8 {
    tw_t w[7];   // node nd | 2nd, 2nd+1 | 4nd .. 4nd+3
};
template <int SKIP>
__device__ __forceinline__ void load_tw8(Tw8 &t, const tw_t *__restrict__ tw, unsigned nd) {
    if (SKIP < 1) t.w[0] = __ldg(tw + nd);
    if (SKIP < 2) {
        t.w[1] = __ldg(tw + 2 * nd);
        t.w[2] = __ldg(tw + 2 * nd + 1);
    }
#pragma unroll
    for (int q = 0; q < 4; q++) t.w[3 + q] = __ldg(tw + 4 * nd + q);
}
template <int SKIP>
__device__ __forceinline__ void fwd8(u64 (&x)[8], const Tw8 &t, const ModConst &m) {
    if (SKIP < 1) {
#pragma unroll
        for (int e = 0; e < 4; e++) ct_bfly(x[e], x[e + 4], t.w[0].x, t.w[0].y, m);
    }
    if (SKIP < 2) {
        ct_bfly(x[0], x[2], t.w[1].x, t.w[1].y, m);
        ct_bfly(x[1], x[3], t.w[1].x, t.w[1].y, m);
        ct_bfly(x[4], x[6], t.w[2].x, t.w[2].y, m);
        ct_bfly(x[5], x[7], t.w[2].x, t.w[2].y, m);
    }
#pragma unroll
    for (int q = 0; q < 4; q++) ct_bfly(x[2 * q], x[2 * q + 1], t.w[3 + q].x, t.w[3 + q].y, m);
}
template <int SKIP>
__device__ __forceinline__ void inv8(u64 (&x)[8], const Tw8 &t, const ModConst &m) {
#pragma unroll
    for (int q = 0; q < 4; q++) gs_bfly(x[2 * q], x[2 * q + 1], t.w[3 + q].x, t.w[3 + q].y, m);
    if (SKIP < 2) {
        gs_bfly(x[0], x[2], t.w[1].x, t.w[1].y, m);
        gs_bfly(x[1], x[3], t.w[1].x, t.w[1].y, m);
        gs_bfly(x[4], x[6], t.w[2].x, t.w[2].y, m);
        gs_bfly(x[5], x[7], t.w[2].x, t.w[2].y, m);
    }
    if (SKIP < 1) {
#pragma unroll
        for (int e = 0; e < 4; e++) gs_bfly(x[e], x[e + 4], t.w[0].x, t.w[0].y, m);
    }
}
// convenience: load + apply
template <int SKIP>
__device__ __forceinline__ void fwd8(u64 (&x)[8], const tw_t *__restrict__ tw, unsigned nd, const ModConst &m) {
    Tw8 t;
    load_tw8<SKIP>(t, tw, nd);
    fwd8<SKIP>(x, t, m);
}
template <int SKIP>
__device__ __forceinline__ void inv8(u64 (&x)[8], const tw_t *__restrict__ tw, unsigned nd, const ModConst &m) {
    Tw8 t;
    load_tw8<SKIP>(t, tw, nd);
    inv8<SKIP>(x, t, m);
}

// ---- column pass -------------------------------------------------------------------------------
// tile = 64 rows x 32 columns starting at column c0; warp k (0..7), lane = column offset.
// element e <-> row  (k + 8e)  on the coarse side   (stride 8 rows)
//               row  (8k + e)  on the fine side     (stride 1 row)
template <int LOGN>
__device__ __forceinline__ int col_coarse_idx(int c0, int e) {
    int k = ntt_tid() >> 5, lane = ntt_tid() & 31;
    return (k + 8 * e) * NttGeo<LOGN>::N2 + c0 + lane;
}
template <int LOGN>
__device__ __forceinline__ int col_fine_idx(int c0, int e) {
    int k = ntt_tid() >> 5, lane = ntt_tid() & 31;
    return (8 * k + e) * NttGeo<LOGN>::N2 + c0 + lane;
}

// forward: x holds the coarse-side elements (values < 8p); returns fine-side elements, lazy (< 8p + 2^32)
template <int LOGN>
__device__ __forceinline__ void fwd_col_pass(u64 (&x)[8], const tw_t *__restrict__ tw, const ModConst &m, u64 *smem) {
    const int k = ntt_tid() >> 5, lane = ntt_tid() & 31;
    fwd8<0>(x, tw, 1u, m);  // stages 0..2, root node
    Tw8 t2;
    load_tw8<0>(t2, tw, 8u + k);   // in flight across the exchange
#pragma unroll
    for (int e = 0; e < 8; e++) smem[(k + 8 * e) * 32 + lane] = x[e];
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 8; e++) x[e] = smem[(8 * k + e) * 32 + lane];
    fwd8<0>(x, t2, m);  // stages 3..5
}

// inverse: x holds fine-side elements in [0,2p); returns coarse-side elements, scaled by N^-1,
// in [0,4p)
template <int LOGN>
__device__ __forceinline__ void inv_col_pass(u64 (&x)[8], const tw_t *__restrict__ twi, const ModConst &m, u64 *smem) {
    const int k = ntt_tid() >> 5, lane = ntt_tid() & 31;
    inv8<0>(x, twi, 8u + k, m);  // stages 5..3
    Tw8 t2;
    load_tw8<1>(t2, twi, 1u);
#pragma unroll
    for (int e = 0; e < 8; e++) smem[(8 * k + e) * 32 + lane] = x[e];
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 8; e++) x[e] = smem[(k + 8 * e) * 32 + lane];
    inv8<1>(x, t2, m);  // stages 2..1
    // stage 0 with N^-1 folded into both outputs
#pragma unroll
    for (int e = 0; e < 4; e++) {
        u64 s = x[e] + x[e + 4];
        u64 d = x[e] + m.gsc - x[e + 4];
        x[e] = shoup_lazy(s, m.ninv, m.ninvs, m.negp);
        x[e + 4] = shoup_lazy(d, m.w1ni, m.w1nis, m.negp);
    }
}

// ---- row pass ---------------------------------------------------------------------------------
// tile = 2048 contiguous words starting at global index t0 (a whole number of rows).
// strided side : element e <-> local index rr*N2 + k + T*e       (rr = tid / T, k = tid % T)
// contiguous   : element e <-> local index 8*tid + e
template <int LOGN>
__device__ __forceinline__ int row_strided_li(int e) {
    typedef NttGeo<LOGN> G;
    int rr = ntt_tid() / G::T, k = ntt_tid() % G::T;
    return rr * G::N2 + k + G::T * e;
}
__device__ __forceinline__ int row_contig_li(int e) { return 8 * (int)ntt_tid() + e; }

template <int LOGN>
__device__ __forceinline__ int row_mid_li(int e) {
    typedef NttGeo<LOGN> G;
    constexpr int T8 = G::T / 8;  // stride of the middle radix step
    int rr = ntt_tid() / G::T, k = ntt_tid() % G::T;
    int a = k / T8, k2 = k % T8;
    return rr * G::N2 + a * G::T + k2 + T8 * e;
}

// forward: x = strided-side elements (lazy) -> contiguous-side elements, lazy (< 8p + 2^32)
template <int LOGN>
__device__ __forceinline__ void fwd_row_pass(u64 (&x)[8], const tw_t *__restrict__ tw, const ModConst &m, int t0, u64 *smem) {
    typedef NttGeo<LOGN> G;
    constexpr int SK3 = 3 - (G::REM > 0 ? G::REM : 3);
    Tw8 ta, tb;
    // stages 6..8
    load_tw8<0>(ta, tw, 64u + ((unsigned)(t0 + row_strided_li<LOGN>(0)) >> (LOGN - 6)));
    fwd8<0>(x, ta, m);
    load_tw8<0>(tb, tw, 512u + ((unsigned)(t0 + row_mid_li<LOGN>(0)) >> (LOGN - 9)));   // next step, in flight
#pragma unroll
    for (int e = 0; e < 8; e++) smem[sw(row_strided_li<LOGN>(e))] = x[e];
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 8; e++) x[e] = smem[sw(row_mid_li<LOGN>(e))];
    // stages 9..11
    fwd8<0>(x, tb, m);
    if (G::REM > 0) {
        load_tw8<SK3>(ta, tw, (1u << (LOGN - 3)) + ((unsigned)(t0 + row_contig_li(0)) >> 3));
        // each thread overwrites exactly the slots it has just read: no barrier needed before
#pragma unroll
        for (int e = 0; e < 8; e++) smem[sw(row_mid_li<LOGN>(e))] = x[e];
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 8; e++) x[e] = smem[sw(row_contig_li(e))];
        // last REM stages; virtual root stage is n-3
        fwd8<SK3>(x, ta, m);
    }
}

// inverse: x = contiguous-side elements in [0,2p) -> strided-side elements, lazy (< 4p + 2^47)
template <int LOGN>
__device__ __forceinline__ void inv_row_pass(u64 (&x)[8], const tw_t *__restrict__ twi, const ModConst &m, int t0, u64 *smem) {
    typedef NttGeo<LOGN> G;
    constexpr int SK3 = 3 - (G::REM > 0 ? G::REM : 3);
    Tw8 ta, tb;
    load_tw8<0>(tb, twi, 512u + ((unsigned)(t0 + row_mid_li<LOGN>(0)) >> (LOGN - 9)));
    if (G::REM > 0) {
        load_tw8<SK3>(ta, twi, (1u << (LOGN - 3)) + ((unsigned)(t0 + row_contig_li(0)) >> 3));
        inv8<SK3>(x, ta, m);
#pragma unroll
        for (int e = 0; e < 8; e++) smem[sw(row_contig_li(e))] = x[e];
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 8; e++) x[e] = smem[sw(row_mid_li<LOGN>(e))];
    }
    inv8<0>(x, tb, m);
    load_tw8<0>(ta, twi, 64u + ((unsigned)(t0 + row_strided_li<LOGN>(0)) >> (LOGN - 6)));
#pragma unroll
    for (int e = 0; e < 8; e++) smem[sw(row_mid_li<LOGN>(e))] = x[e];   // same slots this thread read
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 8; e++) x[e] = smem[sw(row_strided_li<LOGN>(e))];
    inv8<0>(x, ta, m);
}

// ================================================================================================
// FP64 versions of the same passes (primes below 2^41, see modarith.cuh).  Twiddles are plain
// doubles (8 bytes per node).  Shared memory holds the doubles' bit patterns.
struct Tw8d {
    double w[7];
};
template <int SKIP>
__device__ __forceinline__ void load_tw8d(Tw8d &t, const double *__restrict__ tw, unsigned nd) {
    if (SKIP < 1) t.w[0] = __ldg(tw + nd);
    if (SKIP < 2) {
        t.w[1] = __ldg(tw + 2 * nd);
        t.w[2] = __ldg(tw + 2 * nd + 1);
    }
#pragma unroll
    for (int q = 0; q < 4; q++) t.w[3 + q] = __ldg(tw + 4 * nd + q);
}
template <int SKIP>
__device__ __forceinline__ void fwd8d(double (&x)[8], const Tw8d &t, const FpConst &f) {
    if (SKIP < 1) {
#pragma unroll
        for (int e = 0; e < 4; e++) ct_bfly_fp(x[e], x[e + 4], t.w[0], f);
    }
    if (SKIP < 2) {
        ct_bfly_fp(x[0], x[2], t.w[1], f);
        ct_bfly_fp(x[1], x[3], t.w[1], f);
        ct_bfly_fp(x[4], x[6], t.w[2], f);
        ct_bfly_fp(x[5], x[7], t.w[2], f);
    }
#pragma unroll
    for (int q = 0; q < 4; q++) ct_bfly_fp(x[2 * q], x[2 * q + 1], t.w[3 + q], f);
}
template <int SKIP>
__device__ __forceinline__ void inv8d(double (&x)[8], const Tw8d &t, const FpConst &f) {
#pragma unroll
    for (int q = 0; q < 4; q++) gs_bfly_fp(x[2 * q], x[2 * q + 1], t.w[3 + q], f);
    if (SKIP < 2) {
        gs_bfly_fp(x[0], x[2], t.w[1], f);
        gs_bfly_fp(x[1], x[3], t.w[1], f);
        gs_bfly_fp(x[4], x[6], t.w[2], f);
        gs_bfly_fp(x[5], x[7], t.w[2], f);
    }
    if (SKIP < 1) {
#pragma unroll
        for (int e = 0; e < 4; e++) gs_bfly_fp(x[e], x[e + 4], t.w[0], f);
    }
}

// forward column pass: |x| grows by at most 2p per stage (12p over the pass)
template <int LOGN>
__device__ __forceinline__ void fwd_col_pass_fp(double (&x)[8], const double *__restrict__ tw, const FpConst &f, double *smem) {
    const int k = ntt_tid() >> 5, lane = ntt_tid() & 31;
    Tw8d t1, t2;
    load_tw8d<0>(t1, tw, 1u);
    load_tw8d<0>(t2, tw, 8u + k);
    fwd8d<0>(x, t1, f);
#pragma unroll
    for (int e = 0; e < 8; e++) smem[(k + 8 * e) * 32 + lane] = x[e];
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 8; e++) x[e] = smem[(8 * k + e) * 32 + lane];
    fwd8d<0>(x, t2, f);
}
// inverse column pass: inputs reduced (|x| < p), returns values scaled by N^-1 with |x| < 2p
template <int LOGN>
__device__ __forceinline__ void inv_col_pass_fp(double (&x)[8], const double *__restrict__ twi, const FpConst &f, double *smem) {
    const int k = ntt_tid() >> 5, lane = ntt_tid() & 31;
    Tw8d t1, t2;
    load_tw8d<0>(t1, twi, 8u + k);
    load_tw8d<1>(t2, twi, 1u);
    inv8d<0>(x, t1, f);
#pragma unroll
    for (int e = 0; e < 8; e++) smem[(8 * k + e) * 32 + lane] = x[e];
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 8; e++) x[e] = smem[(k + 8 * e) * 32 + lane];
    inv8d<1>(x, t2, f);
#pragma unroll
    for (int e = 0; e < 4; e++) {
        double s = __dadd_rn(x[e], x[e + 4]);
        double d = __dadd_rn(x[e], -x[e + 4]);
        x[e] = fp_mulmod(s, f.ninv, f);
        x[e + 4] = fp_mulmod(d, f.w1ni, f);
    }
}
// forward row pass: strided-side in, contiguous-side out
template <int LOGN>
__device__ __forceinline__ void fwd_row_pass_fp(double (&x)[8], const double *__restrict__ tw, const FpConst &f, int t0, double *smem) {
    typedef NttGeo<LOGN> G;
    constexpr int SK3 = 3 - (G::REM > 0 ? G::REM : 3);
    Tw8d ta, tb;
    load_tw8d<0>(ta, tw, 64u + ((unsigned)(t0 + row_strided_li<LOGN>(0)) >> (LOGN - 6)));
    load_tw8d<0>(tb, tw, 512u + ((unsigned)(t0 + row_mid_li<LOGN>(0)) >> (LOGN - 9)));
    fwd8d<0>(x, ta, f);
#pragma unroll
    for (int e = 0; e < 8; e++) smem[sw(row_strided_li<LOGN>(e))] = x[e];
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 8; e++) x[e] = smem[sw(row_mid_li<LOGN>(e))];
    fwd8d<0>(x, tb, f);
    if (G::REM > 0) {
        load_tw8d<SK3>(ta, tw, (1u << (LOGN - 3)) + ((unsigned)(t0 + row_contig_li(0)) >> 3));
#pragma unroll
        for (int e = 0; e < 8; e++) smem[sw(row_mid_li<LOGN>(e))] = x[e];
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 8; e++) x[e] = smem[sw(row_contig_li(e))];
        fwd8d<SK3>(x, ta, f);
    }
}
// inverse row pass: contiguous-side in (|x| <= p), strided-side out, reduced to |x| < p at the end
template <int LOGN>
__device__ __forceinline__ void inv_row_pass_fp(double (&x)[8], const double *__restrict__ twi, const FpConst &f, int t0, double *smem) {
    typedef NttGeo<LOGN> G;
    constexpr int SK3 = 3 - (G::REM > 0 ? G::REM : 3);
    Tw8d ta, tb;
    load_tw8d<0>(tb, twi, 512u + ((unsigned)(t0 + row_mid_li<LOGN>(0)) >> (LOGN - 9)));
    if (G::REM > 0) {
        load_tw8d<SK3>(ta, twi, (1u << (LOGN - 3)) + ((unsigned)(t0 + row_contig_li(0)) >> 3));
        inv8d<SK3>(x, ta, f);
#pragma unroll
        for (int e = 0; e < 8; e++) smem[sw(row_contig_li(e))] = x[e];
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 8; e++) x[e] = smem[sw(row_mid_li<LOGN>(e))];
    }
    inv8d<0>(x, tb, f);
    load_tw8d<0>(ta, twi, 64u + ((unsigned)(t0 + row_strided_li<LOGN>(0)) >> (LOGN - 6)));
#pragma unroll
    for (int e = 0; e < 8; e++) smem[sw(row_mid_li<LOGN>(e))] = x[e];
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 8; e++) x[e] = smem[sw(row_strided_li<LOGN>(e))];
    inv8d<0>(x, ta, f);
#pragma unroll
    for (int e = 0; e < 8; e++) x[e] = fp_reduce(x[e], f);   // sums doubled 9 times: bring back below p
}
