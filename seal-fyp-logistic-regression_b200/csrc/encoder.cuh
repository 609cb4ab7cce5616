// encoder.cuh -- CKKSEncoder::encode / decode on the device.
//
// Reference call sites: encode(vector<double>, scale, plain) linear_transformation2.cpp:328-330,
// logistic_regression_ckks.cpp:225,305,590-611; encode(double, scale, plain)
// logistic_regression_ckks.cpp:78,157,331; decode(plain, vector<double>&)
// linear_transformation2.cpp:384, logistic_regression_ckks.cpp:365,499.
//
// Slot i of a plaintext m(X) is m(zeta^(3^i)), zeta = exp(2 pi i / 2N); the other N/2 odd powers
// carry the conjugates.  With k_i = (3^i mod 2N - 1)/2 the embedding is an ordinary length-N DFT
// wrapped in a twist by zeta^-j:
//     encode:  v[k_i] = z_i, v[N-1-k_i] = conj(z_i);  c_j = round(scale/N * Re(DFT_-(v)_j * zeta^-j))
//     decode:  v_j = c_j/scale * zeta^j;              z_i = DFT_+(v)[k_i]
// The DFT is a decimation-in-frequency radix-2 transform (natural order in, bit-reversed order out,
// the consumer indexes with bitrev) in two kernels: the top log2(N)-11 stages in registers over
// elements 2048 apart (coalesced across threads), then 11 stages on 2048-point tiles in shared
// memory with a 1024-entry root table built per CTA.  Complex doubles throughout: results agree
// with SEAL's encoder within floating-point tolerance (coefficients within +-1), not bit-exactly --
// north_star compares encode/decode by tolerance.
#pragma once
#include "kernels.cuh"

constexpr int FFT_TILE = 2048;

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// v[b][k_i] = z_i, v[b][N-1-k_i] = conj(z_i) = z_i for real input; slots >= count are zero.  Every
// position of v is written exactly once (the 3^i and their negatives cover the odd residues mod 2N).
__global__ void k_enc_scatter(const double *values, int count, const uint32_t *kidx, double2 *v, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (i >= n / 2) return;
    const double z = i < count ? values[(size_t)b * count + i] : 0.0;
    const uint32_t k = kidx[i];
    v[(size_t)b * n + k] = make_double2(z, 0.0);
    v[(size_t)b * n + (n - 1 - k)] = make_double2(z, 0.0);
}

// top stages: thread = one column c of the 2048-wide matrix, R = N/2048 elements in registers
template <int LOGN, int SGN>
__global__ void __launch_bounds__(256) k_fft_top(double2 *v) {
    constexpr int N = 1 << LOGN, R = N / FFT_TILE;
    const int c = blockIdx.x * 256 + threadIdx.x;
    double2 *p = v + (size_t)blockIdx.y * N + c;
    double2 x[R];
#pragma unroll
    for (int r = 0; r < R; r++) x[r] = p[(size_t)r * FFT_TILE];
#pragma unroll
    for (int hr = R / 2; hr >= 1; hr >>= 1) {
#pragma unroll
        for (int r = 0; r < R; r++) {
            if (r & hr) continue;
            const int j = c + (r & (2 * hr - 1)) * FFT_TILE;       // index inside the block of 2h = 2*hr*2048
            double sn, cs;
            sincospi((double)SGN * (double)j / (double)(hr * FFT_TILE), &sn, &cs);
            const double2 a = x[r], bq = x[r + hr];
            x[r] = make_double2(a.x + bq.x, a.y + bq.y);
            x[r + hr] = cmul(make_double2(a.x - bq.x, a.y - bq.y), make_double2(cs, sn));
        }
    }
#pragma unroll
    for (int r = 0; r < R; r++) p[(size_t)r * FFT_TILE] = x[r];
}

// the remaining 11 stages on one 2048-point tile in shared memory
template <int SGN>
__global__ void __launch_bounds__(256) k_fft_tile(double2 *v, int n) {
    __shared__ double2 x[FFT_TILE];        // 32 KB
    __shared__ double2 root[FFT_TILE / 2]; // 16 KB: exp(SGN * 2 pi i * k / 2048)
    double2 *p = v + (size_t)blockIdx.y * n + (size_t)blockIdx.x * FFT_TILE;
    for (int k = threadIdx.x; k < FFT_TILE / 2; k += 256) {
        double sn, cs;
        sincospi((double)SGN * (double)k / (double)(FFT_TILE / 2), &sn, &cs);
        root[k] = make_double2(cs, sn);
    }
    for (int k = threadIdx.x; k < FFT_TILE; k += 256) x[k] = p[k];
    __syncthreads();
    for (int h = FFT_TILE / 2; h >= 1; h >>= 1) {
        const int stride = (FFT_TILE / 2) / h;
        for (int q = threadIdx.x; q < FFT_TILE / 2; q += 256) {
            const int j = q & (h - 1), lo = ((q - j) << 1) + j;
            const double2 a = x[lo], bq = x[lo + h];
            x[lo] = make_double2(a.x + bq.x, a.y + bq.y);
            x[lo + h] = cmul(make_double2(a.x - bq.x, a.y - bq.y), root[j * stride]);
        }
        __syncthreads();
    }
    for (int k = threadIdx.x; k < FFT_TILE; k += 256) p[k] = x[k];
}

// integer-valued double -> residue mod p.  |r| < 2^62 goes through int64; larger values are
// split into a 53-bit mantissa and a power of two (SEAL decomposes them into 64-bit words)
__device__ __forceinline__ u64 residue_of(double r, const ModConst &m) {
    const bool neg = r < 0.0;
    const double a = fabs(r);
    u64 res;
    if (a < 4611686018427387904.0) {
        res = (u64)a % m.p;
    } else {
        int e;
        const double fr = frexp(a, &e);                 // a = fr * 2^e, fr in [0.5, 1)
        const u64 mant = (u64)ldexp(fr, 53);            // exact
        e -= 53;
        u64 pw = 1, base = 2;
        for (; e > 0; e >>= 1) {
            if (e & 1) pw = mulmod(pw, base, m);
            base = mulmod(base, base, m);
        }
        res = mulmod(mant % m.p, pw, m);
    }
    return (neg && res) ? m.p - res : res;
}

// c_j = round(scale/N * Re(V[bitrev(j)] * zeta^-j)) reduced into every limb (coefficient form)
template <int LOGN>
__global__ void k_enc_round(const double2 *v, double scale_over_n, DView out, int limbs, Tables t) {
    constexpr int N = 1 << LOGN;
    const int j = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    const double2 val = v[(size_t)b * N + (__brev((unsigned)j) >> (32 - LOGN))];
    double sn, cs;
    sincospi((double)j / (double)N, &sn, &cs);
    const double r = rint((val.x * cs + val.y * sn) * scale_over_n);
    u64 *o = out.data + (size_t)b * out.bs + j;
    for (int l = 0; l < limbs; l++) o[(size_t)l * N] = residue_of(r, t.mod[l]);
}

struct ConstResidues {
    u64 r[32];
};
// encode(double): the plaintext is a constant polynomial, whose NTT is that constant in every position
__global__ void k_enc_fill(DView out, ConstResidues cr, int n) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, l = blockIdx.y, b = blockIdx.z;
    if (j < n) out.data[(size_t)b * out.bs + (size_t)l * n + j] = cr.r[l];
}

struct HalfDigits {
    u64 d[32];   // mixed-radix digits of (Q_L - 1)/2 for the level being decoded
};
// CRT-compose coefficient j (Garner mixed-radix digits, exact), centre it on (-Q/2, Q/2], convert
// to double, divide by scale and twist: v_j = c_j/scale * zeta^j
template <int LOGN>
__global__ void k_dec_compose(const u64 *res, int limbs, HalfDigits half, double inv_scale, double2 *v, Tables t) {
    constexpr int N = 1 << LOGN;
    const int j = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    const u64 *x = res + (size_t)b * limbs * N + j;
    u64 d[32];
    for (int i = 0; i < limbs; i++) {
        const ModConst m = t.mod[i];
        u64 u = x[(size_t)i * N];
        for (int k = 0; k < i; k++) {
            const u64 dk = reduce64(d[k], m);
            u = mulmod(submod(u, dk, m.p), t.inv[k * t.K + i], m);
        }
        d[i] = u;
    }
    int cmp = 0;   // sign of (x - half) decided at the most significant differing digit
    for (int i = limbs - 1; i >= 0 && cmp == 0; i--) cmp = d[i] > half.d[i] ? 1 : (d[i] < half.d[i] ? -1 : 0);
    const bool neg = cmp > 0;
    if (neg) {     // Q - x = (Q - 1 - x) + 1: complement every digit, then increment
        bool carry = true;
        for (int i = 0; i < limbs; i++) {
            u64 q = t.mod[i].p, c = q - 1 - d[i];
            if (carry) {
                c += 1;
                carry = c == q;
                if (carry) c = 0;
            }
            d[i] = c;
        }
    }
    double acc = 0.0;
    for (int i = limbs - 1; i >= 0; i--) acc = acc * (double)t.mod[i].p + (double)d[i];
    if (neg) acc = -acc;
    acc *= inv_scale;
    double sn, cs;
    sincospi((double)j / (double)N, &sn, &cs);
    v[(size_t)b * N + j] = make_double2(acc * cs, acc * sn);
}

// z_i = Re(V[bitrev(k_i)])
template <int LOGN>
__global__ void k_dec_gather(const double2 *v, const uint32_t *kidx, double *values) {
    constexpr int N = 1 << LOGN;
    const int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (i >= N / 2) return;
    values[(size_t)b * (N / 2) + i] = v[(size_t)b * N + (__brev(kidx[i]) >> (32 - LOGN))].x;
}

// ------------------------------------------------------------------------------------ sampling
// KeyGenerator / Encryptor randomness on the device: SEAL's sample_poly_ternary, sample_poly_normal
// (sigma 3.2, clipped at 6 sigma) and sample_poly_uniform.  The generator is ChaCha20 (RFC 8439 block
// function, 20 rounds) under a 256-bit key supplied by the caller -- the seal/seal.h shim and client.py draw
// that key from the operating system's CSPRNG -- with the block counter = (coefficient, polynomial | limb |
// attempt) and the nonce = the caller's stream id: a cryptographic PRF in counter mode, reproducible and
// order-independent.  (SEAL 3.4.5's default generator is likewise a CSPRNG: Blake2-based.)
struct SampleKey {
    uint32_t k[8];
};
__device__ __forceinline__ uint32_t rotl32(uint32_t v, int n) { return __funnelshift_l(v, v, n); }
#define CHACHA_QR(a, b, c, d) \
    a += b; d ^= a; d = rotl32(d, 16); c += d; b ^= c; b = rotl32(b, 12); a += b; d ^= a; d = rotl32(d, 8); c += d; b ^= c; b = rotl32(b, 7);
// first four output words of the ChaCha20 block (key, counter = {c0, c1}, nonce = {n0, n1})
__device__ __forceinline__ uint4 chacha20_block4(const SampleKey &key, uint32_t c0, uint32_t c1, uint32_t n0, uint32_t n1) {
    uint32_t x0 = 0x61707865u, x1 = 0x3320646eu, x2 = 0x79622d32u, x3 = 0x6b206574u;
    uint32_t x4 = key.k[0], x5 = key.k[1], x6 = key.k[2], x7 = key.k[3], x8 = key.k[4], x9 = key.k[5], x10 = key.k[6], x11 = key.k[7];
    uint32_t x12 = c0, x13 = c1, x14 = n0, x15 = n1;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        CHACHA_QR(x0, x4, x8, x12) CHACHA_QR(x1, x5, x9, x13) CHACHA_QR(x2, x6, x10, x14) CHACHA_QR(x3, x7, x11, x15)
        CHACHA_QR(x0, x5, x10, x15) CHACHA_QR(x1, x6, x11, x12) CHACHA_QR(x2, x7, x8, x13) CHACHA_QR(x3, x4, x9, x14)
    }
    return make_uint4(x0 + 0x61707865u, x1 + 0x3320646eu, x2 + 0x79622d32u, x3 + 0x6b206574u);
}

// kind 0: ternary {-1,0,1}; kind 1: rounded normal, sigma 3.2, |v| <= 19.  The same small integer is
// reduced into every limb (coefficient form; the caller transforms).
__global__ void k_sample_small(DView out, int limbs, int n, int kind, SampleKey key, u64 stream_id, Tables t) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (j >= n) return;
    const uint4 r = chacha20_block4(key, (unsigned)j, (unsigned)b, (unsigned)stream_id, (unsigned)(stream_id >> 32));
    int v;
    if (kind == 0) {
        // uniform on {-1, 0, 1}: multiply-shift of a 64-bit draw (bias < 2^-62)
        v = (int)__umul64hi(((u64)r.x << 32) | r.y, 3) - 1;
    } else {
        const double u1 = ((double)(((u64)r.x << 32) | r.y) + 0.5) * 5.421010862427522e-20;   // (0, 1]
        const double u2 = ((double)r.z + 0.5) * 2.3283064365386963e-10;
        double sn, cs;
        sincospi(2.0 * u2, &sn, &cs);
        const double z = 3.2 * sqrt(-2.0 * log(u1)) * cs;
        v = (int)rint(fmin(fmax(z, -19.0), 19.0));
    }
    u64 *o = out.data + (size_t)b * out.bs + j;
    for (int l = 0; l < limbs; l++) {
        const u64 p = t.mod[l].p;
        o[(size_t)l * n] = v >= 0 ? (u64)v : p - (u64)(-v);
    }
}

// uniform residues in [0, q_l) for every limb by Lemire's multiply-shift with rejection: x uniform on
// 64 bits, result hi64(x * q); draws whose low product word falls below 2^64 mod q are redrawn, which
// makes the result exactly uniform (rejection probability q / 2^64 < 1/16 per draw).
__global__ void k_sample_uniform(DView out, int limbs, int n, SampleKey key, u64 stream_id, Tables t) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (j >= n) return;
    u64 *o = out.data + (size_t)b * out.bs + j;
    for (int l = 0; l < limbs; l++) {
        const u64 p = t.mod[l].p;
        const u64 thresh = (0 - p) % p;       // 2^64 mod p
        u64 res = 0;
        for (unsigned attempt = 0; attempt < 8; attempt++) {
            // counter word 1: polynomial (20 bits) | limb (6 bits) | attempt (3 bits) | domain bit for "uniform"
            const uint4 r = chacha20_block4(key, (unsigned)j, (unsigned)b | ((unsigned)l << 20) | (attempt << 26) | 0x80000000u,
                                            (unsigned)stream_id, (unsigned)(stream_id >> 32));
            const u64 x = ((u64)r.x << 32) | r.y;
            res = __umul64hi(x, p);
            if (x * p >= thresh) break;        // unbiased draw (rejects with probability p / 2^64 < 1/16)
        }
        o[(size_t)l * n] = res;
    }
}
