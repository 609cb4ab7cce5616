// epoch_driver.cpp -- the headline workload (BASELINE.json config 5) driven from C++: one encrypted LR training
// epoch over C = 8 features x R samples in mini-batches of B (column layout), N = 32768, {60, 40 x 8, 60},
// degree-7 tree sigmoid, through b200::column_epoch_gradient / apply_gradient (ckks_b200_lr.h).  Prints the
// decrypted new weights against plaintext LR and the measured epochs/s.
// usage: epoch_driver [R = 32768] [B = 8192] [timed epochs = 2]
#include <chrono>
#include <iostream>
#include <random>

#include "ckks_b200_lr.h"

using namespace std;
using namespace seal;

int main(int argc, char **argv) {
    const int R = argc > 1 ? atoi(argv[1]) : 32768, B = argc > 2 ? atoi(argv[2]) : 8192, reps = argc > 3 ? atoi(argv[3]) : 2;
    const int C = 8, degree = 7;
    const double lr = 0.1;
    EncryptionParameters params(scheme_type::CKKS);
    size_t n = 32768;
    params.set_poly_modulus_degree(n);
    params.set_coeff_modulus(CoeffModulus::Create(n, {60, 40, 40, 40, 40, 40, 40, 40, 40, 60}));
    auto context = SEALContext::Create(params);
    KeyGenerator keygen(context);
    PublicKey pk = keygen.public_key();
    SecretKey sk = keygen.secret_key();
    RelinKeys rk = keygen.relin_keys();
    GaloisKeys gk = keygen.galois_keys();
    Encryptor encryptor(context, pk);
    Evaluator evaluator(context);
    Decryptor decryptor(context, sk);
    CKKSEncoder encoder(context);
    const double scale = pow(2.0, 40);
    const size_t slots = encoder.slot_count();

    mt19937_64 rng(10);
    normal_distribution<double> gauss(0.0, 1.0);
    uniform_real_distribution<double> uni(0.0, 1.0);
    vector<vector<double>> X(R, vector<double>(C));
    vector<double> y(R), w(C), wtrue(C);
    for (auto &v : wtrue) v = 2 * uni(rng) - 1;
    for (auto &v : w) v = 2 * uni(rng) - 1;
    for (int i = 0; i < R; i++) {
        double z = 0;
        for (int j = 0; j < C; j++) X[i][j] = gauss(rng), z += X[i][j] * wtrue[j];
        y[i] = 1.0 / (1.0 + exp(-z)) > uni(rng) ? 1.0 : 0.0;
    }
    b200::ColumnLayout lay(R, C, B, slots);
    auto enc = [&](const vector<double> &v) {
        Plaintext p;
        encoder.encode(v, scale, p);
        Ciphertext c;
        encryptor.encrypt(p, c);
        return c;
    };
    vector<Ciphertext> cols(lay.M * C), labels(lay.M), wb(C);
    for (int m = 0; m < lay.M; m++) {
        for (int j = 0; j < C; j++) cols[m * C + j] = enc(lay.column(X, m, j));
        labels[m] = enc(lay.labels(y, m));
    }
    for (int j = 0; j < C; j++) wb[j] = enc(vector<double>(slots, w[j]));
    vector<double> wvec(slots, 0.0);
    for (int j = 0; j < C; j++) wvec[j] = w[j];
    Ciphertext wct = enc(wvec);

    auto &eng = *wct.poly().eng;
    auto epoch = [&] {
        Ciphertext grad = b200::column_epoch_gradient(cols, labels, wb, B, scale, evaluator, encoder, gk, rk, encryptor, params, degree, true);
        return b200::apply_gradient(grad, wct, lr, R, scale, evaluator, encoder);
    };
    Ciphertext neww = epoch();          // warm-up (graphs, pools, plans)
    ckks_stream_sync(eng.ctx, nullptr);
    auto t0 = chrono::high_resolution_clock::now();
    for (int i = 0; i < reps; i++) neww = epoch();
    ckks_stream_sync(eng.ctx, nullptr);
    double s = chrono::duration<double>(chrono::high_resolution_clock::now() - t0).count() / reps;

    Plaintext p;
    vector<double> got;
    decryptor.decrypt(neww, p);
    encoder.decode(p, got);
    vector<double> coeffs = b200::sigmoid_coeffs(degree), g(C, 0.0);
    for (int i = 0; i < R; i++) {
        double z = 0, sg = 0, pw = 1;
        for (int j = 0; j < C; j++) z += X[i][j] * w[j];
        for (double c : coeffs) sg += c * pw, pw *= z;
        for (int j = 0; j < C; j++) g[j] += X[i][j] * (sg - y[i]);
    }
    double err = 0;
    for (int j = 0; j < C; j++) err = max(err, fabs(got[j] - (w[j] - lr / R * g[j])));
    cout << "C++ epoch: " << C << " features x " << R << " samples, mini-batches of " << B << ": " << s << " s per epoch = " << 1.0 / s
         << " epochs/s; max |decrypt - plaintext LR| = " << err << endl;
    cout << (err < 1e-3 ? "EPOCH OK" : "FAILED") << endl;
    return err < 1e-3 ? 0 : 1;
}
