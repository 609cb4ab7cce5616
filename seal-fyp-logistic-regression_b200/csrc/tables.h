// tables.h -- host-side constant tables for the CKKS engine (see tables.cpp).
#pragma once
#include <cstdint>
#include <vector>

namespace ckks {

struct HostTables {
    std::vector<uint64_t> mod;      // [K][12] {p, 2p, ratio_lo, ratio_hi, N^-1, shoup, w1*N^-1, shoup, gsc, 4p, 2^64-p, hi32(4p)}
    std::vector<uint64_t> twf;      // [K][N][2] forward twiddle tree {w, shoup(w)}
    std::vector<uint64_t> twi;      // [K][N][2] inverse twiddle tree
    std::vector<uint64_t> inv;      // [K][K]  q_a^-1 mod q_j
    std::vector<uint64_t> invs;     // [K][K]  Shoup companions
    std::vector<uint64_t> halfmod;  // [K][K]  (q_a >> 1) mod q_j
    // FP64 butterfly path for primes below 2^41 (see modarith.cuh): per prime {p, 1/p, N^-1, w1*N^-1, ok}
    std::vector<double> fpc;        // [K][6]  (ok stored as 1.0 / 0.0, then 2^31 mod p)
    std::vector<double> twfd;       // [K][N]  forward twiddle tree as doubles (zero for large primes)
    std::vector<double> twid;       // [K][N]  inverse twiddle tree as doubles
};

bool is_prime_u64(uint64_t n);
// throws std::invalid_argument on unusable primes
void build_tables(int log_n, const std::vector<uint64_t> &primes, HostTables &out);
// index table of the NTT-domain Galois automorphism: out[i] = in[perm[i]]
void build_galois_perm(int log_n, uint64_t galois_elt, std::vector<uint32_t> &perm);
// SEAL steps_to_galois_elt; 0 when |steps| >= N/2
uint64_t galois_elt_from_step(int log_n, int steps);
// SEAL naf(): terms least-significant first
std::vector<int> naf_terms(int steps);

}  // namespace ckks
