// Exercises the seal/seal.h shim the way the reference's drivers do (SEAL 3.4 AND 3.6 spellings):
// parameters -> keys -> encode/encrypt -> the evaluator calls of Linear_Transform_Plain,
// cipher_dot_product and a Horner step -> decrypt/decode -> compare with plaintext math.
// Prints "OK" and exits 0 when every check is within tolerance.
#include <cmath>
#include <cstdio>
#include <iostream>
#include <sstream>
#include <vector>

#include "seal/seal.h"

using namespace std;
using namespace seal;

// Linear_Transform_Plain exactly as the reference writes it (helper.h:237-262)
static Ciphertext lt_plain(Ciphertext ct, vector<Plaintext> U_diagonals, GaloisKeys gal_keys, EncryptionParameters params) {
    SEALContext context(params);
    Evaluator evaluator(context);
    Ciphertext ct_rot;
    evaluator.rotate_vector(ct, -U_diagonals.size(), gal_keys, ct_rot);
    Ciphertext ct_new;
    evaluator.add(ct, ct_rot, ct_new);
    vector<Ciphertext> ct_result(U_diagonals.size());
    evaluator.multiply_plain(ct_new, U_diagonals[0], ct_result[0]);
    for (int l = 1; l < (int)U_diagonals.size(); l++) {
        Ciphertext temp_rot;
        evaluator.rotate_vector(ct_new, l, gal_keys, temp_rot);
        evaluator.multiply_plain(temp_rot, U_diagonals[l], ct_result[l]);
    }
    Ciphertext ct_prime;
    evaluator.add_many(ct_result, ct_prime);
    return ct_prime;
}

static double max_err(const vector<double> &got, const vector<double> &want) {
    double e = 0;
    for (size_t i = 0; i < want.size(); i++) e = max(e, fabs(got[i] - want[i]));
    return e;
}

int main() {
    // --- 3.4-style setup (linear_transformation2.cpp:229-239)
    EncryptionParameters params(scheme_type::CKKS);
    size_t N = 8192;
    params.set_poly_modulus_degree(N);
    params.set_coeff_modulus(CoeffModulus::Create(N, {60, 40, 40, 60}));
    auto context = SEALContext::Create(params);
    KeyGenerator keygen(context);
    PublicKey pk = keygen.public_key();
    SecretKey sk = keygen.secret_key();
    GaloisKeys gal_keys = keygen.galois_keys();
    RelinKeys relin_keys = keygen.relin_keys();
    Encryptor encryptor(context, pk);
    Evaluator evaluator(context);
    Decryptor decryptor(context, sk);
    CKKSEncoder ckks_encoder(context);
    double scale = pow(2.0, 40);
    if (CoeffModulus::MaxBitCount(N) != 218) return 2;
    if (context->get_context_data(context->first_parms_id())->chain_index() != 2) return 3;
    if (context->key_context_data()->total_coeff_modulus_bit_count() != 200) return 4;

    int d = 10;
    vector<vector<double>> U(d, vector<double>(d));
    vector<double> v(d);
    for (int i = 0; i < d; i++) {
        v[i] = (double)(i + 1) / d;
        for (int j = 0; j < d; j++) U[i][j] = (double)((i * 7 + j * 3) % 11) / 11.0;
    }
    vector<Plaintext> diags(d);
    for (int l = 0; l < d; l++) {
        vector<double> dg(d);
        for (int k = 0; k < d; k++) dg[k] = U[k][(k + l) % d];
        ckks_encoder.encode(dg, scale, diags[l]);
    }
    Plaintext pv;
    ckks_encoder.encode(v, scale, pv);
    Ciphertext cv;
    encryptor.encrypt(pv, cv);
    Ciphertext res = lt_plain(cv, diags, gal_keys, params);
    Plaintext pres;
    decryptor.decrypt(res, pres);
    vector<double> out;
    ckks_encoder.decode(pres, out);
    vector<double> want(d, 0.0);
    for (int i = 0; i < d; i++)
        for (int j = 0; j < d; j++) want[i] += U[i][j] * v[j];
    double e1 = max_err(out, want);
    printf("linear transform max err %.3g\n", e1);
    if (!(e1 < 1e-4)) return 10;

    // --- 3.6-style setup (logistic_regression_ckks.cpp:418-441) on the same parameters
    EncryptionParameters parms2(scheme_type::ckks);
    parms2.set_poly_modulus_degree(N);
    parms2.set_coeff_modulus(CoeffModulus::Create(N, {60, 40, 40, 60}));
    SEALContext context2(parms2);
    KeyGenerator keygen2(context2);
    PublicKey pk2;
    keygen2.create_public_key(pk2);
    RelinKeys rk2;
    keygen2.create_relin_keys(rk2);
    GaloisKeys gk2;
    keygen2.create_galois_keys(gk2);
    Encryptor enc2(context2, pk2);
    Evaluator ev2(context2);
    Decryptor dec2(context2, keygen2.secret_key());
    CKKSEncoder encd2(context2);

    // cipher_dot_product sequence (helper.h:416-502)
    int size = 8;
    vector<double> a(size), b(size);
    double dot = 0;
    for (int i = 0; i < size; i++) {
        a[i] = 0.1 * (i + 1);
        b[i] = 1.0 - 0.05 * i;
        dot += a[i] * b[i];
    }
    Plaintext pa, pb;
    encd2.encode(a, scale, pa);
    encd2.encode(b, scale, pb);
    Ciphertext ca, cb, mult;
    enc2.encrypt(pa, ca);
    enc2.encrypt(pb, cb);
    ev2.multiply(ca, cb, mult);
    ev2.relinearize_inplace(mult, rk2);
    ev2.rescale_to_next_inplace(mult);
    Ciphertext zero_filled, dup;
    ev2.rotate_vector(mult, -size, gk2, zero_filled);
    ev2.add(mult, zero_filled, dup);
    for (int i = 1; i < size; i++) {
        ev2.rotate_vector_inplace(dup, 1, gk2);
        ev2.add_inplace(mult, dup);
    }
    mult.scale() = pow(2, (int)log2(mult.scale()));
    Plaintext pm;
    dec2.decrypt(mult, pm);
    vector<double> dm;
    encd2.decode(pm, dm);
    printf("dot product %.6f (expected %.6f)\n", dm[0], dot);
    if (!(fabs(dm[0] - dot) < 1e-3)) return 11;

    // Horner step (logistic_regression_ckks.cpp:174-198): temp = temp*x, relin, rescale, + coeff
    Plaintext pc1, pc0;
    encd2.encode(0.5, scale, pc1);
    Ciphertext temp;
    enc2.encrypt(pc1, temp);
    ev2.multiply_inplace(temp, ca);
    ev2.relinearize_inplace(temp, rk2);
    ev2.rescale_to_next_inplace(temp);
    encd2.encode(0.25, scale, pc0);
    ev2.mod_switch_to_inplace(pc0, temp.parms_id());
    temp.scale() = pow(2.0, 40);
    ev2.add_plain_inplace(temp, pc0);
    Plaintext pt;
    dec2.decrypt(temp, pt);
    vector<double> dt;
    encd2.decode(pt, dt);
    double e3 = 0;
    for (int i = 0; i < size; i++) e3 = max(e3, fabs(dt[i] - (0.5 * a[i] + 0.25)));
    printf("horner step max err %.3g\n", e3);
    if (!(e3 < 1e-4)) return 12;

    // error behaviour
    bool threw = false;
    try {
        ev2.add_inplace(ca, mult);   // level mismatch
    } catch (const invalid_argument &) {
        threw = true;
    }
    if (!threw) return 13;
    threw = false;
    try {
        Ciphertext r;
        ev2.rotate_vector(ca, (int)N / 2, gk2, r);   // step count too large
    } catch (const invalid_argument &) {
        threw = true;
    }
    if (!threw) return 14;
    // --- SEAL binary streams: save -> load round trips; the reloaded objects must evaluate identically
    {
        stringstream sp, sc, sr, sg, spt;
        params.save(sp);
        EncryptionParameters params2;
        params2.load(sp);
        if (params2.poly_modulus_degree() != N || params2.coeff_modulus().size() != 4 ||
            !(params2.coeff_modulus()[1] == params.coeff_modulus()[1]) || params2.scheme() != scheme_type::CKKS)
            return 15;
        auto context2 = SEALContext::Create(params2);
        vector<double> v{0.5, -0.25, 0.125, 1.0};
        Plaintext pv;
        ckks_encoder.encode(v, scale, pv);
        Ciphertext cv;
        encryptor.encrypt(pv, cv);
        cv.save(sc);
        pv.save(spt);
        relin_keys.save(sr);
        gal_keys.save(sg);
        Ciphertext cv2;
        cv2.load(context2, sc);
        Plaintext pv2;
        pv2.load(context2, spt);
        RelinKeys rk2;
        rk2.load(context2, sr);
        GaloisKeys gk2;
        gk2.load(context2, sg);
        if (cv2.size() != 2 || cv2.coeff_mod_count() != cv.coeff_mod_count() || cv2.scale() != cv.scale()) return 16;
        // same ops on the originals and on the reloaded copies -> identical streams (ciphertext bytes compare equal)
        auto run = [&](Ciphertext c, const Plaintext &p, const RelinKeys &rk, const GaloisKeys &gk) {
            Ciphertext r, q;
            evaluator.rotate_vector(c, 3, gk, r);               // NAF chain through the loaded Galois keys
            evaluator.multiply_plain_inplace(r, p);
            evaluator.rescale_to_next_inplace(r);
            evaluator.multiply(r, r, q);
            evaluator.relinearize_inplace(q, rk);
            stringstream out;
            q.save(out);
            return out.str();
        };
        const string a = run(cv, pv, relin_keys, gal_keys), b = run(cv2, pv2, rk2, gk2);
        if (a.size() < 1000 || a != b) return 17;
        if ((unsigned char)a[0] != 0x5E || (unsigned char)a[1] != 0xA1 || a[2] != 0 || a[3] != 0) return 18;   // SEALHeader
    }
    cout << "OK" << endl;
    return 0;
}
