// Times the shim's per-call paths the reference's programs exercise one object at a time:
// CKKSEncoder::encode / decode, Encryptor::encrypt, Evaluator::rotate_vector / multiply_plain.
// build: g++ -std=c++17 -O2 -I include -I seal-fyp-logistic-regression_b200/include profiles/micro/shim_encode_timing.cpp \
//        seal-fyp-logistic-regression_b200/libckks_b200.so -Wl,-rpath,$PWD/seal-fyp-logistic-regression_b200 -o gpurun_out/shim_timing
#include <chrono>
#include <iostream>

#include "seal/seal.h"
using namespace seal;
using namespace std;

template <class F>
static double us_per_call(int reps, F f) {
    f();
    auto t0 = chrono::high_resolution_clock::now();
    for (int i = 0; i < reps; i++) f();
    auto t1 = chrono::high_resolution_clock::now();
    return chrono::duration<double, micro>(t1 - t0).count() / reps;
}

int main() {
    EncryptionParameters params(scheme_type::CKKS);
    size_t n = 16384;
    params.set_poly_modulus_degree(n);
    params.set_coeff_modulus(CoeffModulus::Create(n, {60, 40, 40, 40, 40, 60}));
    auto context = SEALContext::Create(params);
    KeyGenerator keygen(context);
    PublicKey pk = keygen.public_key();
    SecretKey sk = keygen.secret_key();
    GaloisKeys gk = keygen.galois_keys();
    Encryptor encryptor(context, pk);
    Evaluator evaluator(context);
    Decryptor decryptor(context, sk);
    CKKSEncoder encoder(context);
    double scale = pow(2.0, 40);
    vector<double> v(n / 2);
    for (size_t i = 0; i < v.size(); i++) v[i] = 0.001 * i;
    vector<Plaintext> keep(64);
    int k = 0;
    cout << "encode (new plaintext each call, kept alive): " << us_per_call(64, [&] { encoder.encode(v, scale, keep[k++ % 64]); }) << " us\n";
    Plaintext pt;
    cout << "encode (same destination):                    " << us_per_call(200, [&] { encoder.encode(v, scale, pt); }) << " us\n";
    cout << "encode(double):                               " << us_per_call(200, [&] { encoder.encode(0.5, scale, pt); }) << " us\n";
    encoder.encode(v, scale, pt);
    Ciphertext ct, ct2;
    cout << "encrypt:                                      " << us_per_call(100, [&] { encryptor.encrypt(pt, ct); }) << " us\n";
    cout << "rotate_vector(1):                             " << us_per_call(200, [&] { evaluator.rotate_vector(ct, 1, gk, ct2); }) << " us\n";
    cout << "rotate_vector(5) (NAF 4+1):                   " << us_per_call(200, [&] { evaluator.rotate_vector(ct, 5, gk, ct2); }) << " us\n";
    cout << "multiply_plain:                               " << us_per_call(200, [&] { evaluator.multiply_plain(ct, pt, ct2); }) << " us\n";
    cout << "add_inplace:                                  " << us_per_call(200, [&] { evaluator.add_inplace(ct2, ct2); }) << " us\n";
    Plaintext dec;
    vector<double> out;
    cout << "decrypt + decode:                             " << us_per_call(100, [&] { decryptor.decrypt(ct, dec); encoder.decode(dec, out); }) << " us\n";
    cout << "max |decode - input| = ";
    double err = 0;
    for (size_t i = 0; i < v.size(); i++) err = max(err, fabs(out[i] - v[i]));
    cout << err << "\n";
    return 0;
}
