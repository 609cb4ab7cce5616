/*
 * ckks_oracle.h -- CPU oracle for the CKKS evaluation hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / the timed CPU baseline.
 *
 * PARITY UNPINNED.  The reference (MarwanNour/SEAL-FYP-Logistic-Regression) has
 * no arithmetic of its own: every homomorphic op is a call into Microsoft SEAL
 * (pinned in prose to 3.4.5, reference README.md:6), which is neither vendored
 * in the reference nor installed here.  This file restates SEAL 3.4.5's published
 * CKKS algorithms and anchors them on the reference's own
 * call sites (cited per function).  The reference holds no golden vectors for
 * this path; what can be pinned (SEAL's hard-coded 128-bit default primes, the
 * NAF decomposition examples, plaintext semantics of the reference's self-checks)
 * is pinned in tests/.
 *
 * Data layout everywhere: a polynomial at a level with L limbs is L*N uint64
 * words, limb-major ([limb][coeff]); a ciphertext of size S is S such
 * polynomials back to back ([poly][limb][coeff]); CKKS data is always kept in
 * NTT form (bit-reversed evaluation order), as SEAL does.  Limb j of a level-L
 * object is modulo prime j of the context; the special prime is prime K-1.
 */
#ifndef CKKS_ORACLE_H
#define CKKS_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_ctx orc_ctx;

/* ---- parameter selection (SEAL CoeffModulus::Create / BFVDefault / MaxBitCount) */
int orc_coeff_modulus_create(int log_n, const int *bit_sizes, int count, uint64_t *out);
int orc_bfv_default(int log_n, uint64_t *out, int cap);   /* returns count */
int orc_max_bit_count(int log_n);
int orc_is_prime(uint64_t v);

/* ---- context */
orc_ctx *orc_create(int log_n, int n_primes, const uint64_t *primes);
void orc_destroy(orc_ctx *c);
/* rounding switch for the two divide-by-last-prime steps (both marked unconfirmed there):
 * 0 = floor in both, 1 = round to nearest in both (default), 2 = key-switch mod-down rounds / rescale floors,
 * 3 = key-switch mod-down floors / rescale rounds.  tools/seal_replay.py against real SEAL files decides. */
void orc_set_rounding(orc_ctx *c, int mode);
uint64_t orc_prime(const orc_ctx *c, int j);
uint64_t orc_psi(const orc_ctx *c, int j);      /* minimal primitive 2N-th root */
int orc_log_n(const orc_ctx *c);
int orc_n_primes(const orc_ctx *c);

/* ---- NTT on one limb (prime index j), in place */
void orc_ntt(const orc_ctx *c, int j, uint64_t *a);
void orc_intt(const orc_ctx *c, int j, uint64_t *a);
/* direct O(N^2) evaluation a(psi^(2*bitrev(i)+1)) -- for validating orc_ntt on small N */
void orc_ntt_naive(const orc_ctx *c, int j, const uint64_t *a, uint64_t *out);

/* ---- element-wise evaluator ops on [S][L][N] arrays */
void orc_add(const orc_ctx *c, int S, int L, const uint64_t *a, const uint64_t *b, uint64_t *out);
void orc_sub(const orc_ctx *c, int S, int L, const uint64_t *a, const uint64_t *b, uint64_t *out);
void orc_negate(const orc_ctx *c, int S, int L, const uint64_t *a, uint64_t *out);
/* ct (Sa polys) x ct (Sb polys) -> Sa+Sb-1 polys */
void orc_multiply(const orc_ctx *c, int Sa, int Sb, int L, const uint64_t *a, const uint64_t *b, uint64_t *out);
/* every poly of ct times plain (one poly) */
void orc_multiply_plain(const orc_ctx *c, int S, int L, const uint64_t *ct, const uint64_t *pt, uint64_t *out);
/* plain added to poly 0 (out may alias ct) */
void orc_add_plain(const orc_ctx *c, int S, int L, const uint64_t *ct, const uint64_t *pt, uint64_t *out);
/* 1 if every poly 1..S-1 is identically zero (SEAL "transparent" check) */
int orc_is_transparent(const orc_ctx *c, int S, int L, const uint64_t *ct);

/* ---- key switching
 * key layout: [digit i < K-1][component k < 2][limb j < K][N], NTT form, key level. */
size_t orc_ksk_words(const orc_ctx *c);
/* ct: size-2 ciphertext at level L, updated in place; target: L limbs, NTT form */
void orc_switch_key(const orc_ctx *c, int L, uint64_t *ct, const uint64_t *target, const uint64_t *ksk);
/* size-3 -> size-2; in: [3][L][N], out: [2][L][N] */
void orc_relinearize(const orc_ctx *c, int L, const uint64_t *in, const uint64_t *rlk, uint64_t *out);
/* one Galois automorphism + key switch on a size-2 ciphertext */
void orc_apply_galois(const orc_ctx *c, int L, const uint64_t *in, uint64_t galois_elt,
                      const uint64_t *gk, uint64_t *out);
/* permutation of one NTT-form limb: out[i] = in[index(i)] */
void orc_galois_permute_limb(const orc_ctx *c, uint64_t galois_elt, const uint64_t *in, uint64_t *out);
uint64_t orc_galois_elt_from_step(const orc_ctx *c, int steps);   /* 0 on "step count too large" */
/* NAF of steps, least-significant term first; returns number of terms */
int orc_naf(int steps, int *out, int cap);

/* ---- rescale / mod switch */
/* in: [S][L][N] -> out: [S][L-1][N], divide-and-round by prime L-1 */
void orc_rescale(const orc_ctx *c, int S, int L, const uint64_t *in, uint64_t *out);
/* drop limb L-1: in [S][L][N] -> out [S][L-1][N] */
void orc_mod_switch_drop(const orc_ctx *c, int S, int L, const uint64_t *in, uint64_t *out);

/* ---- keys, encryption, encoding (tolerance-compared side; deterministic from seed) */
/* secret key: [K][N] NTT form, ternary */
void orc_gen_secret(const orc_ctx *c, uint64_t seed, uint64_t *sk);
/* public key: [2][K][N] */
void orc_gen_public(const orc_ctx *c, uint64_t seed, const uint64_t *sk, uint64_t *pk);
/* key-switch key for new secret s' given in NTT form [K][N] */
void orc_gen_ksk(const orc_ctx *c, uint64_t seed, const uint64_t *sk, const uint64_t *new_key, uint64_t *ksk);
void orc_gen_relin_key(const orc_ctx *c, uint64_t seed, const uint64_t *sk, uint64_t *rlk);
void orc_gen_galois_key(const orc_ctx *c, uint64_t seed, const uint64_t *sk, uint64_t galois_elt, uint64_t *gk);
/* pt: [L][N] NTT form; ct out: [2][L][N] */
void orc_encrypt(const orc_ctx *c, uint64_t seed, int L, const uint64_t *pk, const uint64_t *pt, uint64_t *ct);
/* symmetric variant (-(a s + e) + m, a) */
void orc_encrypt_symmetric(const orc_ctx *c, uint64_t seed, int L, const uint64_t *sk, const uint64_t *pt, uint64_t *ct);
/* pt out: [L][N] NTT form = sum_k c_k s^k */
void orc_decrypt(const orc_ctx *c, int S, int L, const uint64_t *sk, const uint64_t *ct, uint64_t *pt);
/* values: n_values doubles (<= N/2), zero padded */
void orc_encode(const orc_ctx *c, int L, const double *values, int n_values, double scale, uint64_t *pt);
void orc_encode_const(const orc_ctx *c, int L, double value, double scale, uint64_t *pt);
/* out: N/2 doubles (real parts) */
void orc_decode(const orc_ctx *c, int L, const uint64_t *pt, double scale, double *out);

#ifdef __cplusplus
}
#endif
#endif
