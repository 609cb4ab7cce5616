// lr_driver.cpp -- one encrypted training iteration with the C++ batched LR functions of ckks_b200_lr.h
// (b200::predict_cipher_weights / update_weights, mirroring logistic_regression_ckks.cpp:208-345 with the
// repairs listed in that header) on synthetic standardised data; the decrypted predictions and updated
// weights are compared with the plaintext computation using the same polynomial sigmoid.
#include <chrono>
#include <iostream>
#include <random>

#include "ckks_b200_lr.h"

using namespace std;
using namespace seal;

int main(int argc, char **argv) {
    const int R = argc > 1 ? atoi(argv[1]) : 64, C = 4, degree = 3;
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

    mt19937_64 rng(5);
    normal_distribution<double> gauss(0.0, 1.0);
    uniform_real_distribution<double> uni(0.0, 1.0);
    vector<vector<double>> X(R, vector<double>(C));
    vector<double> y(R), w(C), wtrue(C);
    for (auto &v : wtrue) v = 2 * uni(rng) - 1;
    for (auto &v : w) v = 4 * uni(rng) - 2;          // logistic_regression_ckks.cpp:552
    for (int i = 0; i < R; i++) {
        double z = 0;
        for (int j = 0; j < C; j++) X[i][j] = gauss(rng), z += X[i][j] * wtrue[j];
        y[i] = 1.0 / (1.0 + exp(-z)) > uni(rng) ? 1.0 : 0.0;
    }
    b200::RowLayout lay(R, C, encoder.slot_count());
    auto enc = [&](const vector<double> &v) {
        Plaintext p;
        encoder.encode(v, scale, p);
        Ciphertext c;
        encryptor.encrypt(p, c);
        return c;
    };
    vector<Ciphertext> rows(R), cols(C);
    for (int i = 0; i < R; i++) rows[i] = enc(lay.row(X, i));
    for (int j = 0; j < C; j++) cols[j] = enc(lay.column(X, j));
    Ciphertext labels = enc(lay.labels(y)), weights = enc(lay.weights(w));

    auto dec = [&](const Ciphertext &c) {
        Plaintext p;
        vector<double> out;
        decryptor.decrypt(c, p);
        encoder.decode(p, out);
        return out;
    };
    vector<double> coeffs = b200::sigmoid_coeffs(degree);
    auto sigma = [&](double x) {
        double s = 0, p = 1;
        for (double c : coeffs) s += c * p, p *= x;
        return s;
    };
    int failures = 0;

    Ciphertext pred = b200::predict_cipher_weights(rows, weights, C, scale, evaluator, encoder, gk, rk, encryptor, params, degree);
    vector<double> got = dec(pred), p_plain(R);
    double err = 0;
    for (int i = 0; i < R; i++) {
        double z = 0;
        for (int j = 0; j < C; j++) z += X[i][j] * w[j];
        p_plain[i] = sigma(z);
        err = max(err, fabs(got[i] - p_plain[i]));
    }
    cout << "predict_cipher_weights: max |decrypt - sigmoid_poly(X w)| = " << err << " over " << R << " rows" << endl;
    if (!(err < 1e-3)) failures++;

    {   // the tree method, degree 7 (the sigmoid of config 5): same rows, deeper polynomial
        vector<double> c7 = b200::sigmoid_coeffs(7);
        Ciphertext p7 = b200::predict_cipher_weights(rows, weights, C, scale, evaluator, encoder, gk, rk, encryptor, params, 7, true);
        vector<double> g7 = dec(p7);
        double e7 = 0;
        for (int i = 0; i < R; i++) {
            double z = 0, s = 0, pw = 1;
            for (int j = 0; j < C; j++) z += X[i][j] * w[j];
            for (double c : c7) s += c * pw, pw *= z;
            e7 = max(e7, fabs(g7[i] - s));
        }
        cout << "predict_cipher_weights (Tree_cipher, degree 7): max error " << e7 << endl;
        if (!(e7 < 1e-3)) failures++;
    }

    auto &eng = *weights.poly().eng;
    ckks_stream_sync(eng.ctx, nullptr);
    auto t0 = chrono::high_resolution_clock::now();
    Ciphertext neww = b200::update_weights(rows, cols, labels, weights, 0.1f, evaluator, encoder, gk, rk, encryptor, scale, params, degree);
    ckks_stream_sync(eng.ctx, nullptr);
    double ms = chrono::duration<double, milli>(chrono::high_resolution_clock::now() - t0).count();
    got = dec(neww);
    err = 0;
    for (int j = 0; j < C; j++) {
        double g = 0;
        for (int i = 0; i < R; i++) g += X[i][j] * (p_plain[i] - y[i]);
        double want = w[j] - 0.1 / R * g;
        err = max(err, fabs(got[j] - want));
    }
    cout << "update_weights: " << ms << " ms, max |decrypt - plaintext LR step| = " << err << " (level "
         << neww.coeff_mod_count() << ")" << endl;
    if (!(err < 1e-3)) failures++;
    {   // train_cipher: three iterations with the weights refreshed by the key holder after each
        Ciphertext trained = b200::train_cipher(rows, cols, labels, weights, 0.1f, 3, R, C, evaluator, encoder, scale, gk, rk, encryptor,
                                                decryptor, params, degree);
        vector<double> wt = dec(trained), wp = w;
        for (int it = 0; it < 3; it++) {
            vector<double> g(C, 0.0);
            for (int i = 0; i < R; i++) {
                double z = 0;
                for (int j = 0; j < C; j++) z += X[i][j] * wp[j];
                double pi = sigma(z) - y[i];
                for (int j = 0; j < C; j++) g[j] += X[i][j] * pi;
            }
            for (int j = 0; j < C; j++) wp[j] -= 0.1 / R * g[j];
        }
        double e3 = 0;
        for (int j = 0; j < C; j++) e3 = max(e3, fabs(wt[j] - wp[j]));
        cout << "train_cipher (3 iterations): max |decrypt - plaintext LR| = " << e3 << endl;
        if (!(e3 < 1e-3)) failures++;
    }
    cout << (failures ? "FAILED" : "LR OK") << endl;
    return failures ? 1 : 0;
}
