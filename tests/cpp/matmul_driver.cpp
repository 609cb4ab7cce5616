// matmul_driver.cpp -- the reference's CC_Matrix_Multiplication (matrix_mult_benchmark.cpp, included
// UNCHANGED with its main() renamed by the preprocessor; compile-time use of the reference tree only)
// against b200::CC_Matrix_Multiplication on identical inputs and keys: bit-identical ciphertexts.
#define main reference_main
#include "matrix_mult_benchmark.cpp"
#undef main
#include "ckks_b200_helper.h"

static std::vector<std::uint64_t> words(const Ciphertext &ct) {
    const auto &p = ct.poly();
    std::vector<std::uint64_t> w, tmp((std::size_t)p.limbs * p.eng->n);
    for (int k = 0; k < p.size; k++) {
        seal::detail::check(ckks_download(p.eng->ctx, tmp.data(), p.buf->p + (std::size_t)k * p.cap * p.eng->n, tmp.size() * 8, nullptr));
        seal::detail::check(ckks_stream_sync(p.eng->ctx, nullptr));
        w.insert(w.end(), tmp.begin(), tmp.end());
    }
    return w;
}

int main() {
    const int d = 4, dd = d * d;
    EncryptionParameters params(scheme_type::CKKS);
    size_t n = 8192;
    params.set_poly_modulus_degree(n);
    params.set_coeff_modulus(CoeffModulus::Create(n, {40, 30, 30, 30, 30, 50}));
    auto context = SEALContext::Create(params);
    KeyGenerator keygen(context);
    PublicKey pk = keygen.public_key();
    SecretKey sk = keygen.secret_key();
    GaloisKeys gk = keygen.galois_keys();
    Encryptor encryptor(context, pk);
    Evaluator evaluator(context);
    Decryptor decryptor(context, sk);
    CKKSEncoder encoder(context);
    double scale = pow(2.0, 30);
    srand(11);
    vector<vector<double>> A(d, vector<double>(d)), B(d, vector<double>(d));
    for (int i = 0; i < d; i++)
        for (int j = 0; j < d; j++) A[i][j] = (double)rand() / RAND_MAX, B[i][j] = (double)rand() / RAND_MAX;
    const double eps = 1e-4;      // large enough to survive rounding at scale 2^30 with 16 of 4096 slots used
    auto encode_set = [&](vector<vector<double>> U) {
        vector<vector<double>> dg = get_all_diagonals(U);
        vector<Plaintext> pts(dd);
        for (int i = 0; i < dd; i++) {
            for (auto &x : dg[i]) x += eps;     // the reference's way around "result ciphertext is transparent"
            encoder.encode(dg[i], scale, pts[i]);
        }
        return pts;
    };
    vector<Plaintext> sigma = encode_set(get_U_sigma(A)), tau = encode_set(get_U_tau(A));
    vector<vector<Plaintext>> V(d - 1), W(d - 1);
    for (int k = 1; k < d; k++) V[k - 1] = encode_set(get_V_k(A, k)), W[k - 1] = encode_set(get_W_k(A, k));
    vector<double> fa(dd), fb(dd);
    for (int i = 0; i < d; i++)
        for (int j = 0; j < d; j++) fa[i * d + j] = A[i][j], fb[i * d + j] = B[i][j];
    Plaintext pa, pb;
    encoder.encode(fa, scale, pa);
    encoder.encode(fb, scale, pb);
    Ciphertext ca, cb;
    encryptor.encrypt(pa, ca);
    encryptor.encrypt(pb, cb);

    auto &eng = *ca.poly().eng;
    auto run = [&](bool batched, Ciphertext &out) {
        ckks_stream_sync(eng.ctx, nullptr);
        auto t0 = chrono::high_resolution_clock::now();
        out = batched ? b200::CC_Matrix_Multiplication(ca, cb, d, sigma, tau, V, W, gk, params)
                      : CC_Matrix_Multiplication(ca, cb, d, sigma, tau, V, W, gk, params);
        ckks_stream_sync(eng.ctx, nullptr);
        return chrono::duration<double, micro>(chrono::high_resolution_clock::now() - t0).count();
    };
    Ciphertext ref, got;
    run(false, ref), run(true, got);              // warm-up
    double t_ref = run(false, ref), t_got = run(true, got);
    bool ok = ref.size() == got.size() && ref.coeff_mod_count() == got.coeff_mod_count() && ref.scale() == got.scale() &&
              words(ref) == words(got);
    cout << "CC_Matrix_Multiplication (d = 4): " << (ok ? "bit-identical" : "MISMATCH") << "   reference sequence " << t_ref
         << " us, batched " << t_got << " us" << endl;
    Plaintext p;
    vector<double> out;
    decryptor.decrypt(got, p);
    encoder.decode(p, out);
    vector<vector<double>> want = test_matrix_mult(A, B, d);
    double err = 0;
    for (int i = 0; i < d; i++)
        for (int j = 0; j < d; j++) err = max(err, fabs(out[i * d + j] - want[i][j]));
    cout << "    max |decrypt - A B| = " << err << endl;
    ok = ok && err < 1e-2;
    cout << (ok ? "ALL BIT-IDENTICAL" : "FAILED") << endl;
    return ok ? 0 : 1;
}
