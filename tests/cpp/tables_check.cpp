// tables_check.cpp -- CPU-only dump of the engine's host-side tables (csrc/tables.cpp: primitive roots, twiddle
// trees, cross-prime inverses, Galois permutations, NAF, step -> Galois element) so that tests/test_params.py can
// compare them with the oracle.  usage: tables_check log_n prime...
#include <cstdio>
#include <cstdlib>

#include "tables.h"

int main(int argc, char **argv) {
    const int log_n = atoi(argv[1]);
    std::vector<uint64_t> primes;
    for (int i = 2; i < argc; i++) primes.push_back(strtoull(argv[i], nullptr, 10));
    ckks::HostTables t;
    ckks::build_tables(log_n, primes, t);
    const size_t n = size_t(1) << log_n, K = primes.size();
    for (size_t j = 0; j < K; j++) {
        // node 1 of the forward tree is psi^(N/2)... the root itself sits at node bitrev(1) = N/2
        printf("psi %zu %llu\n", j, (unsigned long long)t.twf[(j * n + n / 2) * 2]);
        printf("ninv %zu %llu\n", j, (unsigned long long)t.mod[j * 12 + 4]);
        unsigned long long acc = 0;
        for (size_t e = 0; e < n; e++) acc = acc * 1000003ull + t.twf[(j * n + e) * 2] + 7 * t.twi[(j * n + e) * 2];
        printf("twsum %zu %llu\n", j, acc);
    }
    for (size_t a = 0; a < K; a++)
        for (size_t j = 0; j < K; j++)
            if (a != j) printf("inv %zu %zu %llu %llu\n", a, j, (unsigned long long)t.inv[a * K + j], (unsigned long long)t.halfmod[a * K + j]);
    for (int steps : {1, -1, 5, -8, 100, -2000, 0}) {
        uint64_t g = ckks::galois_elt_from_step(log_n, steps);
        printf("elt %d %llu\n", steps, (unsigned long long)g);
        std::vector<uint32_t> perm;
        ckks::build_galois_perm(log_n, g, perm);
        unsigned long long acc = 0;
        for (size_t i = 0; i < n; i++) acc = acc * 1000003ull + perm[i];
        printf("perm %d %llu %u %u %u\n", steps, acc, perm[0], perm[1], perm[n - 1]);
    }
    for (int steps : {1, -1, 3, 5, 7, -8, 11, 100, -2000, 8191, -8192}) {
        printf("naf %d", steps);
        for (int x : ckks::naf_terms(steps)) printf(" %d", x);
        printf("\n");
    }
    return 0;
}
