// tables.cpp -- host-side construction of the per-prime constants the kernels read:
// Barrett ratios, the minimal primitive 2N-th root (SEAL's choice), forward and
// inverse twiddle trees with Shoup companions, and the cross-prime inverses used by
// mod-down / rescale.  Runs once per context; plain C++ (g++), no CUDA.
#include "tables.h"

#include <stdexcept>

namespace ckks {

typedef unsigned __int128 u128;

static uint64_t mul_mod(uint64_t a, uint64_t b, uint64_t p) { return (uint64_t)((u128)a * b % p); }

static uint64_t pow_mod(uint64_t a, uint64_t e, uint64_t p) {
    uint64_t r = 1;
    a %= p;
    for (; e; e >>= 1) {
        if (e & 1) r = mul_mod(r, a, p);
        a = mul_mod(a, a, p);
    }
    return r;
}

static uint64_t inv_mod(uint64_t a, uint64_t p) { return pow_mod(a % p, p - 2, p); }

static uint64_t shoup_of(uint64_t w, uint64_t p) { return (uint64_t)(((u128)w << 64) / p); }

bool is_prime_u64(uint64_t n) {
    if (n < 2) return false;
    const uint64_t witnesses[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    for (uint64_t w : witnesses) {
        if (n == w) return true;
        if (n % w == 0) return false;
    }
    uint64_t d = n - 1;
    int twos = 0;
    while ((d & 1) == 0) {
        d >>= 1;
        ++twos;
    }
    for (uint64_t w : witnesses) {
        uint64_t x = pow_mod(w, d, n);
        if (x == 1 || x == n - 1) continue;
        bool witness_composite = true;
        for (int r = 1; r < twos && witness_composite; ++r) {
            x = mul_mod(x, x, n);
            if (x == n - 1) witness_composite = false;
        }
        if (witness_composite) return false;
    }
    return true;
}

static uint32_t reverse_bits(uint32_t v, int bits) {
    uint32_t r = 0;
    for (int i = 0; i < bits; ++i, v >>= 1) r = (r << 1) | (v & 1);
    return r;
}

// the numerically smallest primitive 2N-th root of unity modulo p
static uint64_t smallest_primitive_root(uint64_t p, uint64_t order) {
    const uint64_t cofactor = (p - 1) / order;
    uint64_t any = 0;
    for (uint64_t base = 2; !any; ++base) {
        uint64_t cand = pow_mod(base, cofactor, p);
        if (pow_mod(cand, order / 2, p) == p - 1) any = cand;
    }
    const uint64_t step = mul_mod(any, any, p);
    uint64_t best = any, walk = any;
    for (uint64_t k = 1; k < order / 2; ++k) {
        walk = mul_mod(walk, step, p);
        if (walk < best) best = walk;
    }
    return best;
}

void build_tables(int log_n, const std::vector<uint64_t> &primes, HostTables &out) {
    const size_t n = size_t(1) << log_n;
    const size_t K = primes.size();
    out.mod.assign(K * 12, 0);
    out.twf.assign(K * n * 2, 0);
    out.twi.assign(K * n * 2, 0);
    out.inv.assign(K * K, 0);
    out.invs.assign(K * K, 0);
    out.halfmod.assign(K * K, 0);
    out.fpc.assign(K * 6, 0.0);
    out.twfd.assign(K * n, 0.0);
    out.twid.assign(K * n, 0.0);
    for (size_t j = 0; j < K; ++j) {
        const uint64_t p = primes[j];
        if ((p >> 60) != 0 || !is_prime_u64(p) || (p - 1) % (2 * n) != 0)
            throw std::invalid_argument("coeff_modulus primes must be at most 60 bits, prime and 1 mod 2N");
        for (size_t i = 0; i < j; ++i)
            if (primes[i] == p) throw std::invalid_argument("coeff_modulus primes must be distinct");
        const u128 ratio = (~(u128)0) / p;
        const uint64_t psi = smallest_primitive_root(p, 2 * n);
        const uint64_t n_inv = inv_mod((uint64_t)n, p);
        uint64_t *tf = &out.twf[j * n * 2], *ti = &out.twi[j * n * 2];
        uint64_t power = 1;
        for (size_t e = 0; e < n; ++e) {  // psi^e lives at tree node bitrev(e)
            const size_t node = reverse_bits((uint32_t)e, log_n);
            tf[2 * node] = power;
            power = mul_mod(power, psi, p);
        }
        for (size_t node = 0; node < n; ++node) {
            tf[2 * node + 1] = shoup_of(tf[2 * node], p);
            ti[2 * node] = inv_mod(tf[2 * node], p);
            ti[2 * node + 1] = shoup_of(ti[2 * node], p);
        }
        uint64_t *m = &out.mod[j * 12];
        m[0] = p;
        m[1] = 2 * p;
        m[2] = (uint64_t)ratio;
        m[3] = (uint64_t)(ratio >> 64);
        m[4] = n_inv;
        m[5] = shoup_of(n_inv, p);
        m[6] = mul_mod(ti[2 * 1], n_inv, p);  // root node of the inverse tree times N^-1
        m[7] = shoup_of(m[6], p);
        const uint64_t need = 4 * p + (uint64_t(1) << 47);
        m[8] = ((need + p - 1) / p) * p;   // multiple of p covering every lazy inverse-butterfly operand
        m[9] = 4 * p;
        m[10] = 0 - p;
        m[11] = (4 * p) >> 32;
        if ((p >> 41) == 0) {   // small prime: exact FP64 butterflies are possible (|values| < 2^51)
            double *f = &out.fpc[j * 6];
            f[0] = (double)p;
            f[1] = 1.0 / (double)p;
            f[2] = (double)n_inv;
            f[3] = (double)m[6];
            f[4] = 1.0;
            f[5] = (double)((uint64_t(1) << 31) % p);
            for (size_t node = 0; node < n; ++node) {
                out.twfd[j * n + node] = (double)tf[2 * node];
                out.twid[j * n + node] = (double)ti[2 * node];
            }
        }
    }
    for (size_t a = 0; a < K; ++a)
        for (size_t j = 0; j < K; ++j) {
            if (a == j) continue;
            const uint64_t q = primes[j];
            out.inv[a * K + j] = inv_mod(primes[a] % q, q);
            out.invs[a * K + j] = shoup_of(out.inv[a * K + j], q);
            out.halfmod[a * K + j] = (primes[a] >> 1) % q;
        }
}

void build_galois_perm(int log_n, uint64_t galois_elt, std::vector<uint32_t> &perm) {
    const uint32_t n = 1u << log_n;
    const uint64_t mask = 2ull * n - 1;
    perm.resize(n);
    for (uint32_t i = 0; i < n; ++i) {
        const uint64_t exponent = (galois_elt * (2ull * reverse_bits(i, log_n) + 1)) & mask;
        perm[i] = reverse_bits((uint32_t)((exponent - 1) >> 1), log_n);
    }
}

uint64_t galois_elt_from_step(int log_n, int steps) {
    const uint64_t n = 1ull << log_n, m = 2 * n;
    if (steps == 0) return m - 1;
    const uint64_t mag = steps < 0 ? (uint64_t)(-(int64_t)steps) : (uint64_t)steps;
    if (mag >= n / 2) return 0;
    const uint64_t exponent = steps < 0 ? n / 2 - mag : mag;
    uint64_t g = 1;
    for (uint64_t i = 0; i < exponent; ++i) g = (g * 3) & (m - 1);
    return g;
}

std::vector<int> naf_terms(int steps) {
    std::vector<int> terms;
    const bool negative = steps < 0;
    int v = negative ? -steps : steps;
    for (int bit = 0; v != 0; ++bit) {
        const int z = (v & 1) ? 2 - (v & 3) : 0;
        v = (v - z) >> 1;
        if (z != 0) terms.push_back((negative ? -z : z) * (1 << bit));
    }
    return terms;
}

}  // namespace ckks
