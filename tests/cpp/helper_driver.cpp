// helper_driver.cpp -- the reference's own helper.h functions (compiled UNCHANGED from the reference tree,
// running one evaluator call at a time through seal/seal.h) against the batched b200:: drop-ins of
// ckks_b200_helper.h, on identical inputs and keys: every returned ciphertext must be bit-identical.
// Built by tests/cpp/build_reference_drivers.sh (needs the reference tree at compile time only).
#include <chrono>
#include <cstring>
#include <iostream>

#include "helper.h"             // reference (global namespace): Linear_Transform_Plain, cipher_dot_product, ...
#include "ckks_b200_helper.h"   // b200::Linear_Transform_Plain, ...

static std::vector<std::uint64_t> words(const Ciphertext &ct) {
    const auto &p = ct.poly();
    std::vector<std::uint64_t> w;
    std::vector<std::uint64_t> tmp((std::size_t)p.limbs * p.eng->n);
    for (int k = 0; k < p.size; k++) {
        seal::detail::check(ckks_download(p.eng->ctx, tmp.data(), p.buf->p + (std::size_t)k * p.cap * p.eng->n, tmp.size() * 8, nullptr));
        seal::detail::check(ckks_stream_sync(p.eng->ctx, nullptr));
        w.insert(w.end(), tmp.begin(), tmp.end());
    }
    return w;
}

static int failures = 0;
static void same(const char *name, const Ciphertext &ref, const Ciphertext &got, double us_ref, double us_got) {
    bool ok = ref.size() == got.size() && ref.coeff_mod_count() == got.coeff_mod_count() && ref.scale() == got.scale() &&
              words(ref) == words(got);
    std::cout << name << ": " << (ok ? "bit-identical" : "MISMATCH") << "   reference sequence " << us_ref << " us, batched "
              << us_got << " us" << std::endl;
    if (!ok) failures++;
}

template <class F>
static double timed(Ciphertext &out, seal::detail::Engine &e, F f) {
    out = f();   // warm-up (plans, pools, lazily loaded kernels)
    ckks_stream_sync(e.ctx, nullptr);
    auto t0 = chrono::high_resolution_clock::now();
    out = f();
    ckks_stream_sync(e.ctx, nullptr);
    return chrono::duration<double, micro>(chrono::high_resolution_clock::now() - t0).count();
}

int main() {
    EncryptionParameters params(scheme_type::CKKS);
    size_t n = 8192;
    params.set_poly_modulus_degree(n);
    params.set_coeff_modulus(CoeffModulus::Create(n, {60, 40, 40, 60}));
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
    double scale = pow(2.0, 40);
    srand(7);
    auto rnd = [] { return (double)rand() / RAND_MAX; };

    const int d = 12;
    vector<vector<double>> U(d, vector<double>(d));
    vector<double> v(d);
    for (auto &row : U)
        for (auto &x : row) x = rnd();
    for (auto &x : v) x = rnd();
    vector<vector<double>> diags = get_all_diagonals(U);
    vector<Plaintext> diag_pt(d);
    vector<Ciphertext> diag_ct(d);
    for (int i = 0; i < d; i++) {
        encoder.encode(diags[i], scale, diag_pt[i]);
        encryptor.encrypt(diag_pt[i], diag_ct[i]);
    }
    Plaintext pv;
    encoder.encode(v, scale, pv);
    Ciphertext cv;
    encryptor.encrypt(pv, cv);
    seal::detail::Engine &eng = *cv.poly().eng;

    Ciphertext ref, got;
    double t_ref, t_got;

    t_ref = timed(ref, eng, [&] { return Linear_Transform_Plain(cv, diag_pt, gk, params); });
    t_got = timed(got, eng, [&] { return b200::Linear_Transform_Plain(cv, diag_pt, gk, params); });
    same("Linear_Transform_Plain (d = 12)", ref, got, t_ref, t_got);
    {   // and it is the matrix-vector product
        Plaintext p;
        vector<double> out;
        decryptor.decrypt(got, p);
        encoder.decode(p, out);
        double err = 0;
        for (int i = 0; i < d; i++) {
            double want = 0;
            for (int j = 0; j < d; j++) want += U[i][j] * v[j];
            err = max(err, fabs(out[i] - want));
        }
        cout << "    max |decrypt - U v| = " << err << endl;
        if (!(err < 1e-4)) failures++;
    }

    {   // SURVEY 8(f4) opt-in mode from C++: hoisted rotations (needs a key per step): same decrypted product, different polynomials
        vector<int> hsteps{-d};
        for (int l = 1; l < d; l++) hsteps.push_back(l);
        GaloisKeys gk_h = keygen.galois_keys(hsteps);
        Ciphertext h;
        double t_h = timed(h, eng, [&] { return b200::Linear_Transform_Plain_hoisted(cv, diag_pt, gk_h, params); });
        Plaintext p;
        vector<double> out;
        decryptor.decrypt(h, p);
        encoder.decode(p, out);
        double err = 0;
        for (int i = 0; i < d; i++) {
            double want = 0;
            for (int j = 0; j < d; j++) want += U[i][j] * v[j];
            err = max(err, fabs(out[i] - want));
        }
        cout << "  b200::Linear_Transform_Plain_hoisted (d = 12): " << t_h << " us, max |decrypt - U v| = " << err << endl;
        if (!(err < 1e-4)) failures++;
    }

    t_ref = timed(ref, eng, [&] { return Linear_Transform_Cipher(cv, diag_ct, gk, evaluator); });
    t_got = timed(got, eng, [&] { return b200::Linear_Transform_Cipher(cv, diag_ct, gk, evaluator); });
    same("Linear_Transform_Cipher (d = 12)", ref, got, t_ref, t_got);

    vector<Plaintext> rot_pt(d);
    for (int i = 0; i < d; i++) {
        vector<double> r(2 * d);
        for (int k = 0; k < 2 * d; k++) r[k] = v[(k + i) % d];
        encoder.encode(r, scale, rot_pt[i]);
    }
    t_ref = timed(ref, eng, [&] { return Linear_Transform_CipherMatrix_PlainVector(rot_pt, diag_ct, gk, evaluator); });
    t_got = timed(got, eng, [&] { return b200::Linear_Transform_CipherMatrix_PlainVector(rot_pt, diag_ct, gk, evaluator); });
    same("Linear_Transform_CipherMatrix_PlainVector", ref, got, t_ref, t_got);

    vector<Ciphertext> rows(d);
    for (int i = 0; i < d; i++) {
        Plaintext p;
        encoder.encode(U[i], scale, p);
        encryptor.encrypt(p, rows[i]);
    }
    t_ref = timed(ref, eng, [&] { return C_Matrix_Encode(rows, gk, evaluator); });
    t_got = timed(got, eng, [&] { return b200::C_Matrix_Encode(rows, gk, evaluator); });
    same("C_Matrix_Encode (12 rows)", ref, got, t_ref, t_got);

    {   // C_Matrix_Decode of the packed matrix: d row ciphertexts, each compared
        ckks_stream_sync(eng.ctx, nullptr);
        auto t0 = chrono::high_resolution_clock::now();
        vector<Ciphertext> r = C_Matrix_Decode(got, d, scale, gk, encoder, evaluator);
        ckks_stream_sync(eng.ctx, nullptr);
        auto t1 = chrono::high_resolution_clock::now();
        vector<Ciphertext> g = b200::C_Matrix_Decode(got, d, scale, gk, encoder, evaluator);
        ckks_stream_sync(eng.ctx, nullptr);
        auto t2 = chrono::high_resolution_clock::now();
        bool ok = r.size() == g.size();
        for (size_t i = 0; ok && i < r.size(); i++) ok = r[i].scale() == g[i].scale() && words(r[i]) == words(g[i]);
        cout << "C_Matrix_Decode (12 rows): " << (ok ? "bit-identical" : "MISMATCH") << "   reference sequence "
             << chrono::duration<double, micro>(t1 - t0).count() << " us, batched " << chrono::duration<double, micro>(t2 - t1).count()
             << " us" << endl;
        if (!ok) failures++;
    }

    const int size = 256;
    vector<double> a(size), b(size);
    double want = 0;
    for (int i = 0; i < size; i++) a[i] = rnd(), b[i] = rnd(), want += a[i] * b[i];
    Plaintext pa, pb;
    encoder.encode(a, scale, pa);
    encoder.encode(b, scale, pb);
    Ciphertext ca, cb;
    encryptor.encrypt(pa, ca);
    encryptor.encrypt(pb, cb);
    t_ref = timed(ref, eng, [&] { return cipher_dot_product(ca, cb, size, rk, gk, evaluator); });
    t_got = timed(got, eng, [&] { return b200::cipher_dot_product(ca, cb, size, rk, gk, evaluator); });
    same("cipher_dot_product (256 slots)", ref, got, t_ref, t_got);
    {
        Plaintext p;
        vector<double> out;
        decryptor.decrypt(got, p);
        encoder.decode(p, out);
        cout << "    dot product " << out[0] << " (expected " << want << ")" << endl;
        if (!(fabs(out[0] - want) < 1e-3)) failures++;
    }

    cout << (failures ? "FAILED" : "ALL BIT-IDENTICAL") << endl;
    return failures ? 1 : 0;
}
