// seal_dump_vectors.cpp -- the SEAL-side half of the cross-check.
//
// Build this against a REAL Microsoft SEAL 3.4.5 install (the version the reference pins, README.md:6) on any machine
// that has one:
//     g++ -O2 -std=c++17 seal_dump_vectors.cpp -o seal_dump_vectors -lseal        (or through SEAL's CMake package)
//     ./seal_dump_vectors outdir [log_n] [n_forty_bit_primes]
// It writes, in SEAL's own binary format (EncryptionParameters / RelinKeys / GaloisKeys / Ciphertext ::save, no
// compression or zlib -- both are read), a parameter set, the keys, input ciphertexts and the ciphertexts SEAL's own
// Evaluator produces for every primitive on this repository's hot path.  `python tools/seal_replay.py outdir` then loads the
// inputs, runs the SAME ops on the CUDA engine (and on the CPU oracle) and compares the evaluated polynomials bit for bit.
// A pass pins the oracle -- and with it every GPU parity test -- to SEAL itself; a fail names the first differing op
// (the candidates are the conventions SURVEY.md Appendix A marks with a warning sign: rounding in mod-down / rescale,
// non-centred digit lift).  Only SEAL 3.4 API is used (scheme_type::CKKS, SEALContext::Create, keygen.relin_keys()).
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "seal/seal.h"

using namespace seal;
using namespace std;

template <class T>
static void dump(const T &obj, const string &path) {
    ofstream f(path, ios::binary);
    obj.save(f);
    if (!f) throw runtime_error("cannot write " + path);
}

int main(int argc, char **argv) {
    if (argc < 2) {
        cerr << "usage: seal_dump_vectors outdir [log_n=13] [forty_bit_primes=2]" << endl;
        return 2;
    }
    const string dir = argv[1];
    const int log_n = argc > 2 ? atoi(argv[2]) : 13;
    const int mid = argc > 3 ? atoi(argv[3]) : 2;
    const size_t N = size_t(1) << log_n;
    EncryptionParameters params(scheme_type::CKKS);
    params.set_poly_modulus_degree(N);
    vector<int> bits{60};
    for (int i = 0; i < mid; i++) bits.push_back(40);
    bits.push_back(60);
    params.set_coeff_modulus(CoeffModulus::Create(N, bits));       // {60, 40 x mid, 60}: the reference's chains
    auto context = SEALContext::Create(params);
    KeyGenerator keygen(context);
    PublicKey pk = keygen.public_key();
    RelinKeys rlk = keygen.relin_keys();
    GaloisKeys gk = keygen.galois_keys(vector<int>{1, -1, 4, -8});  // rotation by 3 = NAF {-1, 4}: two key switches
    Encryptor encryptor(context, pk);
    Evaluator evaluator(context);
    CKKSEncoder encoder(context);
    const double scale = pow(2.0, 40);

    vector<double> x(64), y(64);
    for (int i = 0; i < 64; i++) {
        x[i] = 0.01 * i - 0.3;
        y[i] = 1.0 / (i + 1);
    }
    Plaintext px, py;
    encoder.encode(x, scale, px);
    encoder.encode(y, scale, py);
    Ciphertext cx, cy;
    encryptor.encrypt(px, cx);
    encryptor.encrypt(py, cy);

    dump(params, dir + "/parms.bin");
    dump(rlk, dir + "/relin_keys.bin");
    dump(gk, dir + "/galois_keys.bin");
    dump(cx, dir + "/in_x.ct");
    dump(cy, dir + "/in_y.ct");
    dump(py, dir + "/in_y.pt");

    Ciphertext t, u;
    evaluator.add(cx, cy, t);                       dump(t, dir + "/out_add.ct");
    evaluator.sub(cx, cy, t);                       dump(t, dir + "/out_sub.ct");
    evaluator.multiply_plain(cx, py, t);            dump(t, dir + "/out_multiply_plain.ct");
    evaluator.multiply(cx, cy, t);                  dump(t, dir + "/out_multiply.ct");
    evaluator.relinearize_inplace(t, rlk);          dump(t, dir + "/out_relinearize.ct");
    evaluator.rescale_to_next_inplace(t);           dump(t, dir + "/out_rescale.ct");
    evaluator.rotate_vector(cx, 1, gk, u);          dump(u, dir + "/out_rotate_1.ct");
    evaluator.rotate_vector(cx, 3, gk, u);          dump(u, dir + "/out_rotate_3.ct");
    evaluator.rotate_vector(t, -8, gk, u);          dump(u, dir + "/out_rotate_low_m8.ct");   // one level down
    evaluator.mod_switch_to_next(cx, u);            dump(u, dir + "/out_mod_switch.ct");
    cout << "wrote test vectors to " << dir << " (N = " << N << ", " << bits.size() << " primes)" << endl;
    return 0;
}
