/*
 * ckks_oracle.c -- CPU restatement of the SEAL 3.4.5 CKKS evaluator semantics the
 * reference's hot path runs on.  TEST INFRASTRUCTURE ONLY; PARITY UNPINNED -- see the
 * header of ckks_oracle.h for what that means and why.
 *
 * Each function cites (a) the SEAL 3.4.5 routine whose published algorithm it restates
 * (SEAL is an un-vendored third-party dependency of the reference, pinned in prose at
 * /root/reference/README.md:6) and (b) the reference call sites that reach it.
 * Nothing here is copied from SEAL or from the reference; it is written from
 * SURVEY.md Appendix A.
 *
 * Plain C11 + unsigned __int128, single thread, no dependencies beyond libm.
 */
#define _GNU_SOURCE
#include "ckks_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef uint64_t u64;
typedef unsigned __int128 u128;

/* ------------------------------------------------------------------ modular arithmetic */

typedef struct {
    u64 p;       /* modulus, < 2^62 */
    u64 r0, r1;  /* floor(2^128 / p), low and high word (SEAL SmallModulus::const_ratio) */
    u64 psi;     /* minimal primitive 2N-th root of unity */
    u64 *w, *ws;   /* forward twiddles psi^bitrev(i) and their Shoup companions */
    u64 *wi, *wis; /* inverse twiddles (psi^bitrev(i))^-1 and Shoup companions */
    u64 ninv, ninvs;
} orc_mod;

struct orc_ctx {
    int log_n;
    size_t n;
    int K;
    int round_half;
    orc_mod *m;
    /* CKKS encoder tables */
    size_t *slot_map;     /* [N]: slot i -> coefficient position, i+N/2 -> conjugate position */
    double *root_re, *root_im; /* zeta^bitrev(i), zeta = exp(2 pi i / 2N) */
};

static inline u64 mul_hi(u64 a, u64 b) { return (u64)(((u128)a * b) >> 64); }

static u64 powmod(u64 a, u64 e, u64 p) {
    u64 r = 1 % p;
    a %= p;
    while (e) {
        if (e & 1) r = (u64)((u128)r * a % p);
        a = (u64)((u128)a * a % p);
        e >>= 1;
    }
    return r;
}
static u64 invmod(u64 a, u64 p) { return powmod(a, p - 2, p); }

/* SEAL util::barrett_reduce_128: input (lo,hi) < 2^128, output canonical in [0,p) */
static inline u64 barrett128(u64 lo, u64 hi, const orc_mod *m) {
    u64 carry = mul_hi(lo, m->r0);
    u128 t2 = (u128)lo * m->r1;
    u128 s1 = (u128)(u64)t2 + carry;
    u64 tmp1 = (u64)s1;
    u64 tmp3 = (u64)(t2 >> 64) + (u64)(s1 >> 64);
    t2 = (u128)hi * m->r0;
    s1 = (u128)tmp1 + (u64)t2;
    carry = (u64)(t2 >> 64) + (u64)(s1 >> 64);
    tmp1 = hi * m->r1 + tmp3 + carry;
    tmp3 = lo - tmp1 * m->p;
    return tmp3 >= m->p ? tmp3 - m->p : tmp3;
}
static inline u64 mulmod(u64 a, u64 b, const orc_mod *m) {
    u128 z = (u128)a * b;
    return barrett128((u64)z, (u64)(z >> 64), m);
}
/* SEAL util::barrett_reduce_63 generalised to any 64-bit input */
static inline u64 reduce64(u64 x, const orc_mod *m) {
    u64 q = mul_hi(x, m->r1);
    u64 r = x - q * m->p;
    return r >= m->p ? r - m->p : r;
}
static inline u64 addmod(u64 a, u64 b, u64 p) { u64 s = a + b; return s >= p ? s - p : s; }
static inline u64 submod(u64 a, u64 b, u64 p) { return a >= b ? a - b : a + p - b; }
static inline u64 shoup(u64 w, u64 p) { return (u64)((((u128)w) << 64) / p); }
/* w*x mod p in [0,p) given ws = floor(w 2^64 / p) */
static inline u64 mul_shoup(u64 x, u64 w, u64 ws, u64 p) {
    u64 q = mul_hi(ws, x);
    u64 r = w * x - q * p;
    return r >= p ? r - p : r;
}

/* deterministic Miller-Rabin, exact for all 64-bit inputs */
int orc_is_prime(u64 n) {
    static const u64 bases[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    if (n < 2) return 0;
    for (size_t i = 0; i < sizeof bases / sizeof *bases; i++) {
        if (n == bases[i]) return 1;
        if (n % bases[i] == 0) return 0;
    }
    u64 d = n - 1;
    int r = 0;
    while (!(d & 1)) { d >>= 1; r++; }
    for (size_t i = 0; i < sizeof bases / sizeof *bases; i++) {
        u64 x = powmod(bases[i], d, n);
        if (x == 1 || x == n - 1) continue;
        int comp = 1;
        for (int k = 1; k < r; k++) {
            x = (u64)((u128)x * x % n);
            if (x == n - 1) { comp = 0; break; }
        }
        if (comp) return 0;
    }
    return 1;
}

static size_t bitrev(size_t x, int bits) {
    size_t r = 0;
    for (int i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}

/* ------------------------------------------------------------------ parameter selection */

/* SEAL 3.4.5 util::get_primes (numth.h): walk down from 2^b - 2N + 1 in steps of 2N.
 * Output order: descending. */
static int get_primes(int bit_size, int count, size_t ntt_size, u64 *out) {
    u64 factor = 2 * (u64)ntt_size;
    u64 value = ((u64)1 << bit_size) - factor + 1;
    u64 lower = (u64)1 << (bit_size - 1);
    int got = 0;
    while (got < count && value > lower) {
        if (orc_is_prime(value)) out[got++] = value;
        value -= factor;
    }
    return got == count ? 0 : -1;
}

/* SEAL 3.4.5 CoeffModulus::Create(N, bit_sizes) -- SURVEY A.1.  Reference call sites:
 * linear_transformation2.cpp:229-233, matrix_multiplication.cpp:144-147,
 * logistic_regression_ckks.cpp:418-424, polynomial.cpp:105-110. */
int orc_coeff_modulus_create(int log_n, const int *bit_sizes, int count, u64 *out) {
    size_t n = (size_t)1 << log_n;
    int mult[64] = {0};
    u64 *lists[64] = {0};
    int rc = 0;
    for (int i = 0; i < count; i++) {
        if (bit_sizes[i] < 2 || bit_sizes[i] > 60) return -1;   /* SEAL_USER_MOD_BIT_COUNT_MAX */
        mult[bit_sizes[i]]++;
    }
    for (int b = 0; b < 64 && !rc; b++) {
        if (!mult[b]) continue;
        lists[b] = malloc(sizeof(u64) * (size_t)mult[b]);
        rc = get_primes(b, mult[b], n, lists[b]);
    }
    if (!rc) {
        /* take from the back of each size's (descending) list */
        for (int i = 0; i < count; i++) out[i] = lists[bit_sizes[i]][--mult[bit_sizes[i]]];
    }
    for (int b = 0; b < 64; b++) free(lists[b]);
    return rc;
}

/* SEAL 3.4.5 util::global_variables::default_coeff_modulus_128 (hard-coded table).
 * Reference call site: benchmark.cpp:137 (CoeffModulus::BFVDefault(N)). */
int orc_bfv_default(int log_n, u64 *out, int cap) {
    static const u64 t12[] = {0xffffee001ULL, 0xffffc4001ULL, 0x1ffffe0001ULL};
    static const u64 t13[] = {0x7fffffd8001ULL, 0x7fffffc8001ULL, 0xfffffffc001ULL, 0xffffff6c001ULL,
                              0xfffffebc001ULL};
    static const u64 t14[] = {0xfffffffd8001ULL,  0xfffffffa0001ULL,  0xfffffff00001ULL,
                              0x1fffffff68001ULL, 0x1fffffff50001ULL, 0x1ffffffee8001ULL,
                              0x1ffffffea0001ULL, 0x1ffffffe88001ULL, 0x1ffffffe48001ULL};
    static const u64 t15[] = {0x7fffffffe90001ULL, 0x7fffffffbf0001ULL, 0x7fffffffbd0001ULL,
                              0x7fffffffba0001ULL, 0x7fffffffaa0001ULL, 0x7fffffffa50001ULL,
                              0x7fffffff9f0001ULL, 0x7fffffff7e0001ULL, 0x7fffffff770001ULL,
                              0x7fffffff380001ULL, 0x7fffffff330001ULL, 0x7fffffff2d0001ULL,
                              0x7fffffff170001ULL, 0x7fffffff150001ULL, 0x7ffffffef00001ULL,
                              0xfffffffff70001ULL};
    const u64 *t;
    int cnt;
    switch (log_n) {
    case 12: t = t12; cnt = 3; break;
    case 13: t = t13; cnt = 5; break;
    case 14: t = t14; cnt = 9; break;
    case 15: t = t15; cnt = 16; break;
    default: return -1;
    }
    if (cnt > cap) return -1;
    memcpy(out, t, sizeof(u64) * (size_t)cnt);
    return cnt;
}

/* SEAL CoeffModulus::MaxBitCount, 128-bit classical security (HE standard table). */
int orc_max_bit_count(int log_n) {
    switch (log_n) {
    case 10: return 27;
    case 11: return 54;
    case 12: return 109;
    case 13: return 218;
    case 14: return 438;
    case 15: return 881;
    default: return 0;
    }
}

/* ------------------------------------------------------------------ context */

/* SEAL util::try_minimal_primitive_root: the numerically smallest primitive 2N-th root.
 * SURVEY A.3. */
static u64 minimal_primitive_root(u64 p, size_t two_n) {
    u64 e = (p - 1) / two_n;
    u64 root = 0;
    for (u64 g = 2; g < p; g++) {
        u64 c = powmod(g, e, p);
        if (powmod(c, two_n / 2, p) == p - 1) { root = c; break; }
    }
    u64 sq = (u64)((u128)root * root % p);
    u64 cur = root, best = root;
    for (size_t i = 0; i < two_n / 2; i++) {   /* all odd powers = all primitive 2N-th roots */
        if (cur < best) best = cur;
        cur = (u64)((u128)cur * sq % p);
    }
    return best;
}

orc_ctx *orc_create(int log_n, int n_primes, const u64 *primes) {
    if (log_n < 2 || log_n > 17 || n_primes < 1) return NULL;
    orc_ctx *c = calloc(1, sizeof *c);
    c->log_n = log_n;
    c->n = (size_t)1 << log_n;
    c->K = n_primes;
    c->round_half = 3;   /* bit 0: key-switch mod-down rounds to nearest, bit 1: rescale does */
    c->m = calloc((size_t)n_primes, sizeof(orc_mod));
    size_t n = c->n;
    for (int j = 0; j < n_primes; j++) {
        orc_mod *m = &c->m[j];
        u64 p = primes[j];
        if (!orc_is_prime(p) || (p - 1) % (2 * n) != 0 || p >> 62) { orc_destroy(c); return NULL; }
        m->p = p;
        u128 ratio = (~(u128)0) / p;  /* p odd => floor((2^128-1)/p) == floor(2^128/p) */
        m->r0 = (u64)ratio;
        m->r1 = (u64)(ratio >> 64);
        m->psi = minimal_primitive_root(p, 2 * n);
        m->w = malloc(sizeof(u64) * n);
        m->ws = malloc(sizeof(u64) * n);
        m->wi = malloc(sizeof(u64) * n);
        m->wis = malloc(sizeof(u64) * n);
        u64 pw = 1;
        for (size_t i = 0; i < n; i++) {   /* psi^i stored at bitrev(i) */
            size_t r = bitrev(i, log_n);
            m->w[r] = pw;
            pw = (u64)((u128)pw * m->psi % p);
        }
        for (size_t i = 0; i < n; i++) {
            m->ws[i] = shoup(m->w[i], p);
            m->wi[i] = invmod(m->w[i], p);
            m->wis[i] = shoup(m->wi[i], p);
        }
        m->ninv = invmod((u64)n % p, p);
        m->ninvs = shoup(m->ninv, p);
    }
    /* CKKS encoder tables (SEAL CKKSEncoder ctor): slot i <-> evaluation point zeta^(3^i) */
    size_t slots = n / 2, two_n = 2 * n;
    c->slot_map = malloc(sizeof(size_t) * n);
    u64 pos = 1;
    for (size_t i = 0; i < slots; i++) {
        c->slot_map[i] = bitrev((size_t)((pos - 1) >> 1), log_n);
        c->slot_map[slots + i] = bitrev((size_t)((two_n - pos - 1) >> 1), log_n);
        pos = pos * 3 % two_n;
    }
    c->root_re = malloc(sizeof(double) * n);
    c->root_im = malloc(sizeof(double) * n);
    for (size_t i = 0; i < n; i++) {
        double ang = 2.0 * M_PI * (double)bitrev(i, log_n) / (double)two_n;
        c->root_re[i] = cos(ang);
        c->root_im[i] = sin(ang);
    }
    return c;
}

void orc_destroy(orc_ctx *c) {
    if (!c) return;
    for (int j = 0; j < c->K; j++) {
        free(c->m[j].w); free(c->m[j].ws); free(c->m[j].wi); free(c->m[j].wis);
    }
    free(c->m); free(c->slot_map); free(c->root_re); free(c->root_im);
    free(c);
}
/* mode: 0 = floor everywhere, 1 = round to nearest everywhere (default), 2 = key switch rounds / rescale floors,
 * 3 = key switch floors / rescale rounds */
void orc_set_rounding(orc_ctx *c, int mode) { c->round_half = mode == 0 ? 0 : mode == 1 ? 3 : mode == 2 ? 1 : 2; }
u64 orc_prime(const orc_ctx *c, int j) { return c->m[j].p; }
u64 orc_psi(const orc_ctx *c, int j) { return c->m[j].psi; }
int orc_log_n(const orc_ctx *c) { return c->log_n; }
int orc_n_primes(const orc_ctx *c) { return c->K; }

/* ------------------------------------------------------------------ NTT */

/* SEAL util::ntt_negacyclic_harvey: Cooley-Tukey, natural in -> bit-reversed out, Harvey lazy
 * butterflies on [0,4p), final correction to [0,p).  SURVEY A.3. */
void orc_ntt(const orc_ctx *c, int j, u64 *a) {
    const orc_mod *m = &c->m[j];
    const u64 p = m->p, two_p = 2 * p;
    size_t n = c->n, t = n >> 1;
    for (size_t mm = 1; mm < n; mm <<= 1, t >>= 1) {
        for (size_t i = 0; i < mm; i++) {
            const u64 W = m->w[mm + i], Ws = m->ws[mm + i];
            u64 *x = a + 2 * i * t, *y = x + t;
            for (size_t k = 0; k < t; k++) {
                u64 X = x[k] - (x[k] >= two_p ? two_p : 0);
                u64 Q = mul_hi(Ws, y[k]);
                u64 T = W * y[k] - Q * p;          /* [0,2p) */
                x[k] = X + T;
                y[k] = X + two_p - T;
            }
        }
    }
    for (size_t k = 0; k < n; k++) {
        u64 v = a[k];
        if (v >= two_p) v -= two_p;
        if (v >= p) v -= p;
        a[k] = v;
    }
}

/* SEAL util::inverse_ntt_negacyclic_harvey: Gentleman-Sande, exact inverse incl. N^-1. */
void orc_intt(const orc_ctx *c, int j, u64 *a) {
    const orc_mod *m = &c->m[j];
    const u64 p = m->p, two_p = 2 * p;
    size_t n = c->n, t = 1;
    for (size_t mm = n >> 1; mm >= 1; mm >>= 1, t <<= 1) {
        for (size_t i = 0; i < mm; i++) {
            const u64 W = m->wi[mm + i], Ws = m->wis[mm + i];
            u64 *x = a + 2 * i * t, *y = x + t;
            for (size_t k = 0; k < t; k++) {
                u64 U = x[k], V = y[k];            /* both in [0,2p) */
                u64 S = U + V;
                x[k] = S - (S >= two_p ? two_p : 0);
                u64 D = U + two_p - V;
                u64 Q = mul_hi(Ws, D);
                y[k] = W * D - Q * p;              /* [0,2p) */
            }
        }
    }
    for (size_t k = 0; k < n; k++) {
        u64 v = a[k];
        if (v >= p) v -= p;
        a[k] = mul_shoup(v, m->ninv, m->ninvs, p);
    }
}

void orc_ntt_naive(const orc_ctx *c, int j, const u64 *a, u64 *out) {
    const orc_mod *m = &c->m[j];
    size_t n = c->n;
    for (size_t i = 0; i < n; i++) {
        u64 x = powmod(m->psi, 2 * bitrev(i, c->log_n) + 1, m->p);
        u64 acc = 0;
        for (size_t k = n; k-- > 0;) acc = addmod(mulmod(acc, x, m), a[k] % m->p, m->p);
        out[i] = acc;
    }
}

/* ------------------------------------------------------------------ element-wise ops */

/* SEAL Evaluator::add_inplace / sub_inplace / negate_inplace (add_poly_poly_coeffmod ...).
 * Reference: helper.h:247,259 (add, add_many), logistic_regression_ckks.cpp:341-342. */
void orc_add(const orc_ctx *c, int S, int L, const u64 *a, const u64 *b, u64 *out) {
    size_t n = c->n;
    for (int s = 0; s < S; s++)
        for (int j = 0; j < L; j++) {
            size_t o = ((size_t)s * L + j) * n;
            u64 p = c->m[j].p;
            for (size_t k = 0; k < n; k++) out[o + k] = addmod(a[o + k], b[o + k], p);
        }
}
void orc_sub(const orc_ctx *c, int S, int L, const u64 *a, const u64 *b, u64 *out) {
    size_t n = c->n;
    for (int s = 0; s < S; s++)
        for (int j = 0; j < L; j++) {
            size_t o = ((size_t)s * L + j) * n;
            u64 p = c->m[j].p;
            for (size_t k = 0; k < n; k++) out[o + k] = submod(a[o + k], b[o + k], p);
        }
}
void orc_negate(const orc_ctx *c, int S, int L, const u64 *a, u64 *out) {
    size_t n = c->n;
    for (int s = 0; s < S; s++)
        for (int j = 0; j < L; j++) {
            size_t o = ((size_t)s * L + j) * n;
            u64 p = c->m[j].p;
            for (size_t k = 0; k < n; k++) out[o + k] = a[o + k] ? p - a[o + k] : 0;
        }
}

/* SEAL Evaluator::ckks_multiply: c_k = sum_{i+j=k} a_i (.) b_j mod q.
 * Reference: helper.h:222,228,432; matrix_multiplication.cpp:105,126. */
void orc_multiply(const orc_ctx *c, int Sa, int Sb, int L, const u64 *a, const u64 *b, u64 *out) {
    size_t n = c->n;
    int So = Sa + Sb - 1;
    u64 *tmp = calloc((size_t)So * L * n, sizeof(u64));
    for (int i = 0; i < Sa; i++)
        for (int k = 0; k < Sb; k++)
            for (int j = 0; j < L; j++) {
                const orc_mod *m = &c->m[j];
                const u64 *x = a + ((size_t)i * L + j) * n, *y = b + ((size_t)k * L + j) * n;
                u64 *z = tmp + ((size_t)(i + k) * L + j) * n;
                for (size_t q = 0; q < n; q++) z[q] = addmod(z[q], mulmod(x[q], y[q], m), m->p);
            }
    memcpy(out, tmp, sizeof(u64) * (size_t)So * L * n);
    free(tmp);
}

/* SEAL Evaluator::multiply_plain_ntt.  Reference: helper.h:250,256,271. */
void orc_multiply_plain(const orc_ctx *c, int S, int L, const u64 *ct, const u64 *pt, u64 *out) {
    size_t n = c->n;
    for (int s = 0; s < S; s++)
        for (int j = 0; j < L; j++) {
            const orc_mod *m = &c->m[j];
            const u64 *x = ct + ((size_t)s * L + j) * n, *y = pt + (size_t)j * n;
            u64 *z = out + ((size_t)s * L + j) * n;
            for (size_t q = 0; q < n; q++) z[q] = mulmod(x[q], y[q], m);
        }
}

/* SEAL Evaluator::add_plain_inplace (CKKS).  Reference: logistic_regression_ckks.cpp:198. */
void orc_add_plain(const orc_ctx *c, int S, int L, const u64 *ct, const u64 *pt, u64 *out) {
    size_t n = c->n;
    if (out != ct) memcpy(out, ct, sizeof(u64) * (size_t)S * L * n);
    for (int j = 0; j < L; j++) {
        u64 p = c->m[j].p;
        for (size_t q = 0; q < n; q++) out[(size_t)j * n + q] = addmod(out[(size_t)j * n + q], pt[(size_t)j * n + q], p);
    }
}

int orc_is_transparent(const orc_ctx *c, int S, int L, const u64 *ct) {
    size_t n = c->n;
    for (size_t q = (size_t)L * n; q < (size_t)S * L * n; q++)
        if (ct[q]) return 0;
    return 1;
}

/* ------------------------------------------------------------------ divide by last prime */

/* The shared tail of SEAL's mod_switch_scale_to_next and switch_key_inplace:
 * given r = INTT(last limb) in [0,qk) (coefficient form), produce for data prime j the NTT of
 *   u_j = ((r + half) mod qk) mod q_j - (half mod q_j)      (rounding; half = qk >> 1)
 * or u_j = r mod q_j when rounding is off.  `work` is one limb of scratch. */
static void last_limb_to(const orc_ctx *c, int jk, int j, const u64 *r, u64 *work, int do_round) {
    const orc_mod *mk = &c->m[jk], *mj = &c->m[j];
    size_t n = c->n;
    u64 half = do_round ? (mk->p >> 1) : 0;
    u64 half_j = reduce64(half, mj);
    for (size_t q = 0; q < n; q++) {
        u64 v = r[q] + half;
        if (v >= mk->p) v -= mk->p;
        work[q] = submod(reduce64(v, mj), half_j, mj->p);
    }
    orc_ntt(c, j, work);
}

/* SEAL Evaluator::mod_switch_scale_to_next (CKKS branch), reached through
 * rescale_to_next_inplace.  Reference: helper.h:441,543; matrix_multiplication.cpp:71-72;
 * logistic_regression_ckks.cpp:189,238,320. */
void orc_rescale(const orc_ctx *c, int S, int L, const u64 *in, u64 *out) {
    size_t n = c->n;
    int jk = L - 1;
    u64 *r = malloc(sizeof(u64) * n), *work = malloc(sizeof(u64) * n);
    for (int s = 0; s < S; s++) {
        memcpy(r, in + ((size_t)s * L + jk) * n, sizeof(u64) * n);
        orc_intt(c, jk, r);
        for (int j = 0; j < L - 1; j++) {
            const orc_mod *mj = &c->m[j];
            u64 inv = invmod(c->m[jk].p % mj->p, mj->p);
            last_limb_to(c, jk, j, r, work, c->round_half & 2);
            const u64 *x = in + ((size_t)s * L + j) * n;
            u64 *z = out + ((size_t)s * (L - 1) + j) * n;
            for (size_t q = 0; q < n; q++) z[q] = mulmod(submod(x[q], work[q], mj->p), inv, mj);
        }
    }
    free(r); free(work);
}

/* SEAL Evaluator::mod_switch_drop_to_next (CKKS ciphertexts and NTT-form plaintexts).
 * Reference: logistic_regression_ckks.cpp:177-182,192,227,286,295; helper.h:536. */
void orc_mod_switch_drop(const orc_ctx *c, int S, int L, const u64 *in, u64 *out) {
    size_t n = c->n;
    for (int s = 0; s < S; s++)
        memmove(out + (size_t)s * (L - 1) * n, in + (size_t)s * L * n, sizeof(u64) * (size_t)(L - 1) * n);
}

/* ------------------------------------------------------------------ key switching */

size_t orc_ksk_words(const orc_ctx *c) { return (size_t)(c->K - 1) * 2 * c->K * c->n; }

static inline const u64 *ksk_limb(const orc_ctx *c, const u64 *ksk, int digit, int comp, int limb) {
    return ksk + (((size_t)digit * 2 + comp) * c->K + limb) * c->n;
}

/* SEAL Evaluator::switch_key_inplace (CKKS branch) -- SURVEY A.7.  Reached from
 * relinearize_inplace (helper.h:440,541; logistic_regression_ckks.cpp:187) and from every
 * Galois step of rotate_vector (helper.h:222,244,255,471,475). */
void orc_switch_key(const orc_ctx *c, int L, u64 *ct, const u64 *target, const u64 *ksk) {
    size_t n = c->n;
    int K = c->K, jP = K - 1;
    int R = L + 1;                                       /* rns_mod_count: data limbs + special */
    u128 *acc = calloc((size_t)2 * R * n, sizeof(u128)); /* [k][j][n] lazy 128-bit sums */
    u64 *d = malloc(sizeof(u64) * n), *t = malloc(sizeof(u64) * n);
    for (int i = 0; i < L; i++) {
        /* digit i: non-negative lift of limb i to the integers */
        memcpy(d, target + (size_t)i * n, sizeof(u64) * n);
        orc_intt(c, i, d);
        for (int jj = 0; jj < R; jj++) {
            int j = jj == L ? jP : jj;                   /* prime index of output limb */
            const u64 *src;
            if (j == i) {
                src = target + (size_t)i * n;            /* already NTT mod q_i */
            } else {
                const orc_mod *mj = &c->m[j];
                for (size_t q = 0; q < n; q++) t[q] = reduce64(d[q], mj);
                orc_ntt(c, j, t);
                src = t;
            }
            for (int k = 0; k < 2; k++) {
                const u64 *key = ksk_limb(c, ksk, i, k, j);
                u128 *a = acc + ((size_t)k * R + jj) * n;
                for (size_t q = 0; q < n; q++) a[q] += (u128)src[q] * key[q];
            }
        }
    }
    /* mod-down by P with rounding, add into (c0, c1) */
    u64 *r = d, *work = t;
    for (int k = 0; k < 2; k++) {
        const orc_mod *mP = &c->m[jP];
        u128 *aP = acc + ((size_t)k * R + L) * n;
        for (size_t q = 0; q < n; q++) r[q] = barrett128((u64)aP[q], (u64)(aP[q] >> 64), mP);
        orc_intt(c, jP, r);
        for (int j = 0; j < L; j++) {
            const orc_mod *mj = &c->m[j];
            u64 pinv = invmod(mP->p % mj->p, mj->p);
            last_limb_to(c, jP, j, r, work, c->round_half & 1);
            u128 *a = acc + ((size_t)k * R + j) * n;
            u64 *z = ct + ((size_t)k * L + j) * n;
            for (size_t q = 0; q < n; q++) {
                u64 v = barrett128((u64)a[q], (u64)(a[q] >> 64), mj);
                v = mulmod(submod(v, work[q], mj->p), pinv, mj);
                z[q] = addmod(z[q], v, mj->p);
            }
        }
    }
    free(acc); free(d); free(t);
}

/* SEAL Evaluator::relinearize_internal for size 3 -> 2: key-switch c2 with rlk[0]. */
void orc_relinearize(const orc_ctx *c, int L, const u64 *in, const u64 *rlk, u64 *out) {
    size_t n = c->n, pw = (size_t)L * n;
    u64 *tgt = malloc(sizeof(u64) * pw);
    memcpy(tgt, in + 2 * pw, sizeof(u64) * pw);
    memmove(out, in, sizeof(u64) * 2 * pw);
    orc_switch_key(c, L, out, tgt, rlk);
    free(tgt);
}

/* SEAL util::steps_to_galois_elt -- SURVEY A.6. */
u64 orc_galois_elt_from_step(const orc_ctx *c, int steps) {
    u64 n = c->n, m = 2 * n;
    if (steps == 0) return m - 1;
    u64 pos = (u64)(steps < 0 ? -(int64_t)steps : steps);
    if (pos >= (n >> 1)) return 0;
    u64 e = steps < 0 ? (n >> 1) - pos : pos;
    u64 g = 1;
    for (u64 i = 0; i < e; i++) g = g * 3 & (m - 1);
    return g;
}

/* SEAL util::naf -- SURVEY A.6.  Terms are emitted least-significant first and
 * Evaluator::rotate_internal applies them in that order. */
int orc_naf(int steps, int *out, int cap) {
    int sign = steps < 0, v = steps < 0 ? -steps : steps, cnt = 0;
    for (int i = 0; v; i++) {
        int z = (v & 1) ? 2 - (v & 3) : 0;
        v = (v - z) >> 1;
        if (z) {
            if (cnt < cap) out[cnt] = (sign ? -z : z) * (1 << i);
            cnt++;
        }
    }
    return cnt;
}

/* SEAL util::apply_galois_ntt: out[i] = in[bitrev(((g (2 bitrev(i)+1) mod 2N) - 1) / 2)]. */
void orc_galois_permute_limb(const orc_ctx *c, u64 g, const u64 *in, u64 *out) {
    size_t n = c->n;
    u64 mask = 2 * n - 1;
    for (size_t i = 0; i < n; i++) {
        u64 raw = (g * (2 * bitrev(i, c->log_n) + 1)) & mask;
        out[i] = in[bitrev((size_t)((raw - 1) >> 1), c->log_n)];
    }
}

/* SEAL Evaluator::apply_galois_inplace (CKKS branch): permute c0 and c1, c1 := 0,
 * key-switch the permuted c1.  Reference: every rotate_vector call (helper.h:222,244,255,318,
 * 350,471,475). */
void orc_apply_galois(const orc_ctx *c, int L, const u64 *in, u64 g, const u64 *gk, u64 *out) {
    size_t n = c->n, pw = (size_t)L * n;
    u64 *tgt = malloc(sizeof(u64) * pw), *c0 = malloc(sizeof(u64) * pw);
    for (int j = 0; j < L; j++) {
        orc_galois_permute_limb(c, g, in + (size_t)j * n, c0 + (size_t)j * n);
        orc_galois_permute_limb(c, g, in + pw + (size_t)j * n, tgt + (size_t)j * n);
    }
    memcpy(out, c0, sizeof(u64) * pw);
    memset(out + pw, 0, sizeof(u64) * pw);
    orc_switch_key(c, L, out, tgt, gk);
    free(tgt); free(c0);
}

/* ------------------------------------------------------------------ sampling */

typedef struct { u64 s[4]; } rng_t;
static u64 splitmix(u64 *x) {
    u64 z = (*x += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
static void rng_seed(rng_t *r, u64 seed) { for (int i = 0; i < 4; i++) r->s[i] = splitmix(&seed); }
static inline u64 rotl(u64 x, int k) { return (x << k) | (x >> (64 - k)); }
static u64 rng_next(rng_t *r) {   /* xoshiro256** */
    u64 *s = r->s, res = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return res;
}
static double rng_unit(rng_t *r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }

/* small signed integer polynomial -> NTT form over primes [0, limbs) */
static void small_poly_to_ntt(const orc_ctx *c, const int *v, int limbs, u64 *out) {
    size_t n = c->n;
    for (int j = 0; j < limbs; j++) {
        u64 p = c->m[j].p;
        u64 *o = out + (size_t)j * n;
        for (size_t q = 0; q < n; q++) o[q] = v[q] >= 0 ? (u64)v[q] : p - (u64)(-v[q]);
        orc_ntt(c, j, o);
    }
}
static void sample_ternary(rng_t *r, size_t n, int *v) {
    for (size_t q = 0; q < n; q++) v[q] = (int)(rng_next(r) % 3) - 1;
}
/* clipped rounded normal, sigma 3.2, |e| <= 6 sigma (SEAL sample_poly_normal) */
static void sample_error(rng_t *r, size_t n, int *v) {
    for (size_t q = 0; q < n; q++) {
        double x;
        do {
            double u1 = rng_unit(r), u2 = rng_unit(r);
            if (u1 < 1e-300) u1 = 1e-300;
            x = 3.2 * sqrt(-2.0 * log(u1)) * cos(2.0 * M_PI * u2);
        } while (fabs(x) > 19.2);
        v[q] = (int)lround(x);
    }
}
static void sample_uniform(const orc_ctx *c, rng_t *r, int limbs, u64 *out) {
    size_t n = c->n;
    for (int j = 0; j < limbs; j++)
        for (size_t q = 0; q < n; q++) out[(size_t)j * n + q] = rng_next(r) % c->m[j].p;
}

/* ------------------------------------------------------------------ keys */

/* SEAL KeyGenerator::generate_sk: ternary secret, NTT form, key level. */
void orc_gen_secret(const orc_ctx *c, u64 seed, u64 *sk) {
    rng_t r; rng_seed(&r, seed);
    int *v = malloc(sizeof(int) * c->n);
    sample_ternary(&r, c->n, v);
    small_poly_to_ntt(c, v, c->K, sk);
    free(v);
}

/* (-(a s + e), a) over primes [0, limbs), NTT form (SEAL util::encrypt_zero_symmetric) */
static void encrypt_zero_sym(const orc_ctx *c, rng_t *r, int limbs, const u64 *sk, u64 *c0, u64 *c1) {
    size_t n = c->n;
    int *e = malloc(sizeof(int) * n);
    u64 *en = malloc(sizeof(u64) * (size_t)limbs * n);
    sample_uniform(c, r, limbs, c1);
    sample_error(r, n, e);
    small_poly_to_ntt(c, e, limbs, en);
    for (int j = 0; j < limbs; j++) {
        const orc_mod *m = &c->m[j];
        for (size_t q = 0; q < n; q++) {
            size_t o = (size_t)j * n + q;
            u64 as = mulmod(c1[o], sk[o], m);
            u64 s = addmod(as, en[o], m->p);
            c0[o] = s ? m->p - s : 0;
        }
    }
    free(e); free(en);
}

void orc_gen_public(const orc_ctx *c, u64 seed, const u64 *sk, u64 *pk) {
    rng_t r; rng_seed(&r, seed);
    encrypt_zero_sym(c, &r, c->K, sk, pk, pk + (size_t)c->K * c->n);
}

/* SEAL KeyGenerator::generate_one_kswitch_key -- SURVEY A.5. */
void orc_gen_ksk(const orc_ctx *c, u64 seed, const u64 *sk, const u64 *new_key, u64 *ksk) {
    rng_t r; rng_seed(&r, seed);
    size_t n = c->n;
    int K = c->K;
    u64 P = c->m[K - 1].p;
    for (int i = 0; i < K - 1; i++) {
        u64 *k0 = ksk + ((size_t)i * 2 + 0) * K * n, *k1 = ksk + ((size_t)i * 2 + 1) * K * n;
        encrypt_zero_sym(c, &r, K, sk, k0, k1);
        const orc_mod *m = &c->m[i];
        u64 factor = P % m->p;
        for (size_t q = 0; q < n; q++) {
            size_t o = (size_t)i * n + q;
            k0[o] = addmod(k0[o], mulmod(new_key[o], factor, m), m->p);
        }
    }
}

void orc_gen_relin_key(const orc_ctx *c, u64 seed, const u64 *sk, u64 *rlk) {
    size_t n = c->n;
    u64 *s2 = malloc(sizeof(u64) * (size_t)c->K * n);
    for (int j = 0; j < c->K; j++)
        for (size_t q = 0; q < n; q++) {
            size_t o = (size_t)j * n + q;
            s2[o] = mulmod(sk[o], sk[o], &c->m[j]);
        }
    orc_gen_ksk(c, seed, sk, s2, rlk);
    free(s2);
}

void orc_gen_galois_key(const orc_ctx *c, u64 seed, const u64 *sk, u64 g, u64 *gk) {
    size_t n = c->n;
    u64 *sg = malloc(sizeof(u64) * (size_t)c->K * n);
    for (int j = 0; j < c->K; j++) orc_galois_permute_limb(c, g, sk + (size_t)j * n, sg + (size_t)j * n);
    orc_gen_ksk(c, seed, sk, sg, gk);
    free(sg);
}

/* ------------------------------------------------------------------ encryption */

/* SEAL Encryptor::encrypt (public key): sample (u pk + e) one level above the target level,
 * divide-and-round by the extra prime, add the plaintext to c0 -- SURVEY A.9. */
void orc_encrypt(const orc_ctx *c, u64 seed, int L, const u64 *pk, const u64 *pt, u64 *ct) {
    rng_t r; rng_seed(&r, seed);
    size_t n = c->n;
    int K = c->K, W = L + 1;   /* primes 0..L: prime L is the next data prime or, at the top level, P */
    int *v = malloc(sizeof(int) * n);
    u64 *u = malloc(sizeof(u64) * (size_t)W * n), *e = malloc(sizeof(u64) * (size_t)W * n);
    u64 *big = malloc(sizeof(u64) * 2 * (size_t)W * n);
    sample_ternary(&r, n, v);
    small_poly_to_ntt(c, v, W, u);
    for (int k = 0; k < 2; k++) {
        sample_error(&r, n, v);
        small_poly_to_ntt(c, v, W, e);
        for (int j = 0; j < W; j++) {
            const orc_mod *m = &c->m[j];
            const u64 *pkl = pk + ((size_t)k * K + j) * n;
            for (size_t q = 0; q < n; q++) {
                size_t o = (size_t)j * n + q;
                big[(size_t)k * W * n + o] = addmod(mulmod(u[o], pkl[q], m), e[o], m->p);
            }
        }
    }
    orc_rescale(c, 2, W, big, ct);
    orc_add_plain(c, 2, L, ct, pt, ct);
    free(v); free(u); free(e); free(big);
}

void orc_encrypt_symmetric(const orc_ctx *c, u64 seed, int L, const u64 *sk, const u64 *pt, u64 *ct) {
    rng_t r; rng_seed(&r, seed);
    encrypt_zero_sym(c, &r, L, sk, ct, ct + (size_t)L * c->n);
    orc_add_plain(c, 2, L, ct, pt, ct);
}

/* SEAL Decryptor::ckks_decrypt: sum_k c_k s^k, NTT form. */
void orc_decrypt(const orc_ctx *c, int S, int L, const u64 *sk, const u64 *ct, u64 *pt) {
    size_t n = c->n;
    for (int j = 0; j < L; j++) {
        const orc_mod *m = &c->m[j];
        for (size_t q = 0; q < n; q++) {
            u64 s = sk[(size_t)j * n + q], acc = ct[((size_t)(S - 1) * L + j) * n + q];
            for (int k = S - 2; k >= 0; k--)
                acc = addmod(mulmod(acc, s, m), ct[((size_t)k * L + j) * n + q], m->p);
            pt[(size_t)j * n + q] = acc;
        }
    }
}

/* ------------------------------------------------------------------ CKKS encoder */

/* complex negacyclic transforms with the same butterfly structure as the NTT (zeta for psi) */
static void cfft_inverse(const orc_ctx *c, double *re, double *im) {
    size_t n = c->n, t = 1;
    for (size_t mm = n >> 1; mm >= 1; mm >>= 1, t <<= 1)
        for (size_t i = 0; i < mm; i++) {
            double wr = c->root_re[mm + i], wi = -c->root_im[mm + i];   /* 1/zeta^k = conj */
            for (size_t k = 2 * i * t; k < 2 * i * t + t; k++) {
                double ur = re[k], ui = im[k], vr = re[k + t], vi = im[k + t];
                re[k] = ur + vr; im[k] = ui + vi;
                double dr = ur - vr, di = ui - vi;
                re[k + t] = dr * wr - di * wi;
                im[k + t] = dr * wi + di * wr;
            }
        }
    for (size_t k = 0; k < n; k++) { re[k] /= (double)n; im[k] /= (double)n; }
}
static void cfft_forward(const orc_ctx *c, double *re, double *im) {
    size_t n = c->n, t = n >> 1;
    for (size_t mm = 1; mm < n; mm <<= 1, t >>= 1)
        for (size_t i = 0; i < mm; i++) {
            double wr = c->root_re[mm + i], wi = c->root_im[mm + i];
            for (size_t k = 2 * i * t; k < 2 * i * t + t; k++) {
                double vr = re[k + t] * wr - im[k + t] * wi, vi = re[k + t] * wi + im[k + t] * wr;
                double ur = re[k], ui = im[k];
                re[k] = ur + vr; im[k] = ui + vi;
                re[k + t] = ur - vr; im[k + t] = ui - vi;
            }
        }
}

/* round(x) reduced into limb j; |x| < 2^126 */
static u64 real_to_limb(double x, const orc_mod *m) {
    double a = fabs(x);
    u64 r;
    if (a < 9.0e18) {
        r = reduce64((u64)llround(a), m);
    } else {
        /* two-word path (SEAL: "max_coeff_bit_count <= 128") */
        double hi_d = floor(a / 18446744073709551616.0);
        double lo_d = a - hi_d * 18446744073709551616.0;
        if (lo_d < 0) lo_d = 0;
        if (lo_d >= 18446744073709551616.0) { lo_d -= 18446744073709551616.0; hi_d += 1; }
        r = barrett128((u64)lo_d, (u64)hi_d, m);
    }
    return (x < 0 && r) ? m->p - r : r;
}

/* SEAL CKKSEncoder::encode(vector<double>, scale, plain) -- SURVEY A.9.
 * Reference: linear_transformation2.cpp:326-331, logistic_regression_ckks.cpp:225,305. */
void orc_encode(const orc_ctx *c, int L, const double *values, int n_values, double scale, u64 *pt) {
    size_t n = c->n, slots = n / 2;
    double *re = calloc(n, sizeof(double)), *im = calloc(n, sizeof(double));
    for (size_t i = 0; i < (size_t)n_values && i < slots; i++) {
        re[c->slot_map[i]] = values[i];
        re[c->slot_map[slots + i]] = values[i];   /* conjugate of a real value */
    }
    cfft_inverse(c, re, im);
    for (int j = 0; j < L; j++) {
        u64 *o = pt + (size_t)j * n;
        for (size_t q = 0; q < n; q++) o[q] = real_to_limb(re[q] * scale, &c->m[j]);
        orc_ntt(c, j, o);
    }
    free(re); free(im);
}

/* SEAL CKKSEncoder::encode(double, scale, plain): constant polynomial round(v*scale);
 * its NTT is the same constant in every position.
 * Reference: logistic_regression_ckks.cpp:78,158,332; polynomial.cpp:139. */
void orc_encode_const(const orc_ctx *c, int L, double value, double scale, u64 *pt) {
    size_t n = c->n;
    for (int j = 0; j < L; j++) {
        u64 v = real_to_limb(value * scale, &c->m[j]);
        for (size_t q = 0; q < n; q++) pt[(size_t)j * n + q] = v;
    }
}

/* SEAL CKKSEncoder::decode: INTT, CRT-compose, centre, scale down, forward complex FFT.
 * CRT composition is done by Garner mixed-radix digits so no multi-precision library is
 * needed. */
void orc_decode(const orc_ctx *c, int L, const u64 *pt, double scale, double *out) {
    size_t n = c->n, slots = n / 2;
    u64 *coef = malloc(sizeof(u64) * (size_t)L * n);
    memcpy(coef, pt, sizeof(u64) * (size_t)L * n);
    for (int j = 0; j < L; j++) orc_intt(c, j, coef + (size_t)j * n);
    /* garner constants: inv[k] = (q_0 ... q_{k-1})^-1 mod q_k */
    u64 inv[64];
    long double radix[64];
    radix[0] = 1.0L;
    for (int k = 0; k < L; k++) {
        u64 prod = 1 % c->m[k].p;
        for (int i = 0; i < k; i++) prod = mulmod(prod, c->m[i].p % c->m[k].p, &c->m[k]);
        inv[k] = invmod(prod, c->m[k].p);
        if (k) radix[k] = radix[k - 1] * (long double)c->m[k - 1].p;
    }
    double *re = calloc(n, sizeof(double)), *im = calloc(n, sizeof(double));
    u64 dg[64];
    for (size_t q = 0; q < n; q++) {
        for (int k = 0; k < L; k++) {
            const orc_mod *m = &c->m[k];
            /* digit k = (x_k - (d0 + d1 q0 + ...)) * inv mod q_k */
            u64 acc = 0, mult = 1 % m->p;
            for (int i = 0; i < k; i++) {
                acc = addmod(acc, mulmod(reduce64(dg[i], m), mult, m), m->p);
                mult = mulmod(mult, c->m[i].p % m->p, m);
            }
            dg[k] = mulmod(submod(coef[(size_t)k * n + q], acc, m->p), inv[k], m);
        }
        int neg = dg[L - 1] >= (c->m[L - 1].p + 1) / 2;
        if (neg) {   /* Q - x: complement every digit, then add one */
            for (int k = 0; k < L; k++) dg[k] = c->m[k].p - 1 - dg[k];
            for (int k = 0; k < L; k++) {
                if (++dg[k] < c->m[k].p) break;
                dg[k] = 0;
            }
        }
        long double v = 0.0L;
        for (int k = L - 1; k >= 0; k--) v += (long double)dg[k] * radix[k];
        re[q] = (double)((neg ? -v : v) / (long double)scale);
    }
    cfft_forward(c, re, im);
    for (size_t i = 0; i < slots; i++) out[i] = re[c->slot_map[i]];
    free(coef); free(re); free(im);
}
