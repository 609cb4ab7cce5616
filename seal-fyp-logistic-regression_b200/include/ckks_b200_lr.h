// ckks_b200_lr.h -- the reference's encrypted logistic-regression functions on the batched engine (C++).
//
// Mirrors logistic_regression_ckks.cpp: Tree_cipher (:55-137), Horner_cipher (:139-205), predict_cipher_weights (:208-266),
// update_weights (:269-345), train_cipher (:348-385) -- same names and argument order, in namespace b200 -- over cipher_dot_product
// (helper.h:416-502).  The per-row / per-feature dot products, which the reference runs one evaluator call at
// a time (R x (C+1) and C x (R+13) key switches), advance in lock-step as batched key switches and the fused
// rotate-and-sum chain; every individual ciphertext still sees the reference's evaluator sequence.
//
// The committed program cannot finish an iteration (rows and weights packed at slot 0 while
// the mask picks slot i; too few primes; scale / level mismatch before the final sub; "CIPHERTEXT IS
// TRANSPARENT" at :336).  The repairs are those of lr.py, client-side only:
//   R1  RowLayout: row i is encoded cyclically at slots i..i+C-1, the weights periodically, so the dot product of
//       row i lands in slot i where the reference's one-hot mask e_i picks it up;
//   R2  the chain needs 9 data primes ({60, 40 x 8, 60} at N = 32768);
//   R3  after the learning-rate multiply the gradient is rescaled, and the weights are brought to its level
//       and scale before sub;
//   R4  train_cipher re-encodes the decoded weights at the top level instead of re-encrypting the exhausted plaintext;
//   R6  the 1/8 input scaling of the sigmoid approximation is folded into the coefficients (sigmoid_coeffs).
#pragma once
#include <algorithm>

#include "ckks_b200_helper.h"

namespace b200 {

// logistic_regression_ckks.cpp:247,251,255 divided by 8^i (R6); zero coefficients are 0.00001 there
inline std::vector<double> sigmoid_coeffs(int degree) {
    std::vector<double> c;
    if (degree == 3) c = {0.5, 1.20069, 0.00001, -0.81562};
    else if (degree == 5) c = {0.5, 1.53048, 0.00001, -2.3533056, 0.00001, 1.3511295};
    else if (degree == 7) c = {0.5, 1.73496, 0.00001, -4.19407, 0.00001, 5.43402, 0.00001, -2.50739};
    else throw std::invalid_argument("Invalid DEGREE");
    double p = 1.0;
    for (auto &x : c) x /= p, p *= 8.0;
    return c;
}

// client-side packing for R samples x C features in N/2 slots (R1): needs R + C <= slots and 2R <= slots
struct RowLayout {
    int R, C;
    std::size_t slots;
    RowLayout(int rows, int cols, std::size_t slot_count) : R(rows), C(cols), slots(slot_count) {
        if ((std::size_t)(R + C) > slots || (std::size_t)(2 * R) > slots) throw std::invalid_argument("R + C and 2R must fit in N/2 slots");
    }
    std::vector<double> row(const std::vector<std::vector<double>> &X, int i) const {
        std::vector<double> v(slots, 0.0);
        for (int s = i; s < i + C; s++) v[s] = X[i][s % C];
        return v;
    }
    std::vector<double> column(const std::vector<std::vector<double>> &X, int j) const {
        std::vector<double> v(slots, 0.0);
        for (int i = 0; i < R; i++) v[i] = X[i][j];
        return v;
    }
    std::vector<double> weights(const std::vector<double> &w) const {
        std::vector<double> v(slots, 0.0);
        for (int s = 0; s < R + C; s++) v[s] = w[s % C];
        return v;
    }
    std::vector<double> labels(const std::vector<double> &y) const {
        std::vector<double> v(slots, 0.0);
        for (int i = 0; i < R; i++) v[i] = y[i];
        return v;
    }
};

// logistic_regression_ckks.cpp:139-205 -- (((a_D) x + a_{D-1}) x + ...) with the reference's level alignment and
// "manual rescale"; one ciphertext, sequential, through the shim's evaluator
inline seal::Ciphertext Horner_cipher(seal::Ciphertext ctx, int degree, const std::vector<double> &coeffs,
                                      seal::CKKSEncoder &ckks_encoder, double scale, seal::Evaluator &evaluator,
                                      seal::Encryptor &encryptor, const seal::RelinKeys &relin_keys, const seal::EncryptionParameters &) {
    if ((int)coeffs.size() != degree + 1) throw std::invalid_argument("coeffs has invalid size");
    std::vector<seal::Plaintext> plain_coeffs(degree + 1);
    for (int i = 0; i <= degree; i++) ckks_encoder.encode(coeffs[i], scale, plain_coeffs[i]);
    seal::Ciphertext temp;
    encryptor.encrypt(plain_coeffs[degree], temp);
    for (int i = degree - 1; i >= 0; i--) {
        if (ctx.coeff_mod_count() > temp.coeff_mod_count()) evaluator.mod_switch_to_inplace(ctx, temp.parms_id());
        else if (ctx.coeff_mod_count() < temp.coeff_mod_count()) evaluator.mod_switch_to_inplace(temp, ctx.parms_id());
        evaluator.multiply_inplace(temp, ctx);
        evaluator.relinearize_inplace(temp, relin_keys);
        evaluator.rescale_to_next_inplace(temp);
        evaluator.mod_switch_to_inplace(plain_coeffs[i], temp.parms_id());
        temp.scale() = std::pow(2.0, 40);
        evaluator.add_plain_inplace(temp, plain_coeffs[i]);
    }
    return temp;
}

// helper.h:505-547 -- x^2..x^degree, each power from the split that minimises multiplicative depth
inline void compute_all_powers(const seal::Ciphertext &ctx, int degree, seal::Evaluator &evaluator, const seal::RelinKeys &relin_keys,
                               std::vector<seal::Ciphertext> &powers) {
    powers.assign(degree + 1, seal::Ciphertext());
    powers[1] = ctx;
    std::vector<int> levels(degree + 1, 0);
    for (int i = 2; i <= degree; i++) {
        int minlevel = i, cand = -1;
        for (int j = 1; j <= i / 2; j++) {
            const int newlevel = std::max(levels[j], levels[i - j]) + 1;
            if (newlevel < minlevel) cand = j, minlevel = newlevel;
        }
        levels[i] = minlevel;
        if (cand < 0) throw std::runtime_error("error");
        seal::Ciphertext temp = powers[cand];
        evaluator.mod_switch_to_inplace(temp, powers[i - cand].parms_id());
        evaluator.multiply(temp, powers[i - cand], powers[i]);
        evaluator.relinearize_inplace(powers[i], relin_keys);
        evaluator.rescale_to_next_inplace(powers[i]);
    }
}

// logistic_regression_ckks.cpp:55-137 -- a_0 + sum_i a_i x^i from the power tree (depth ceil(log2 degree) + 1)
inline seal::Ciphertext Tree_cipher(const seal::Ciphertext &ctx, int degree, double scale, const std::vector<double> &coeffs,
                                    seal::CKKSEncoder &ckks_encoder, seal::Evaluator &evaluator, seal::Encryptor &encryptor,
                                    const seal::RelinKeys &relin_keys, const seal::EncryptionParameters &) {
    if ((int)coeffs.size() != degree + 1) throw std::invalid_argument("coeffs has invalid size");
    std::vector<seal::Plaintext> plain_coeffs(degree + 1);
    for (int i = 0; i <= degree; i++) ckks_encoder.encode(coeffs[i], scale, plain_coeffs[i]);
    std::vector<seal::Ciphertext> powers;
    compute_all_powers(ctx, degree, evaluator, relin_keys, powers);
    seal::Ciphertext enc_result, temp;
    encryptor.encrypt(plain_coeffs[0], enc_result);
    for (int i = 1; i <= degree; i++) {
        evaluator.mod_switch_to_inplace(plain_coeffs[i], powers[i].parms_id());
        evaluator.multiply_plain(powers[i], plain_coeffs[i], temp);
        evaluator.rescale_to_next_inplace(temp);
        evaluator.mod_switch_to_inplace(enc_result, temp.parms_id());
        enc_result.scale() = std::pow(2.0, (int)std::log2(enc_result.scale()));
        temp.scale() = std::pow(2.0, (int)std::log2(enc_result.scale()));
        evaluator.add_inplace(enc_result, temp);
    }
    return enc_result;
}

namespace detail {

// first `limbs` limbs of every object (mod_switch_to_inplace of each, then gather)
template <class T>
inline Batch gather_at(const std::vector<T> &objs, int limbs) {
    if (objs.empty()) throw std::invalid_argument("encrypteds cannot be empty");
    const Poly &p0 = objs[0].poly();
    if (!p0.buf || limbs < 1 || limbs > p0.limbs) throw std::invalid_argument("cannot switch to higher level modulus");
    Batch b(p0.eng, (int)objs.size(), p0.size, limbs, p0.scale);
    const std::size_t n = p0.eng->n;
    for (std::size_t i = 0; i < objs.size(); i++) {
        const Poly &p = objs[i].poly();
        if (!p.buf || p.size != p0.size || p.limbs != p0.limbs) throw std::invalid_argument("encrypted1 and encrypted2 parameter mismatch");
        if (p.scale != p0.scale) throw std::invalid_argument("scale mismatch");
        for (int k = 0; k < p.size; k++)
            check(ckks_copy(p0.eng->ctx, b.buf->p + i * b.entry_words() + (std::size_t)k * limbs * n, p.buf->p + (std::size_t)k * p.cap * n,
                            (std::size_t)limbs * n * 8, nullptr));
    }
    return b;
}

// cipher_dot_product (helper.h:416-502) of every entry of `a` with `b` (one ciphertext for all entries), in lock-step
inline Batch dot_product_batch(const Batch &a, ckks_view vb, double b_scale, int size, const seal::RelinKeys &rk,
                               const seal::GaloisKeys &gk) {
    if (a.size != 2 || vb.size != 2) throw std::invalid_argument("encrypted size must be 2");
    if (a.limbs != vb.limbs || (vb.batch != 1 && vb.batch != a.batch)) throw std::invalid_argument("encrypted1 and encrypted2 parameter mismatch");
    if (a.limbs < 2) throw std::invalid_argument("end of modulus switching chain reached");
    if (!rk.s || !rk.s->keys.count(0)) throw std::invalid_argument("relin_keys is not valid for encryption parameters");
    if (!gk.s || !gk.s->ks) throw std::invalid_argument("galois_keys is not valid for encryption parameters");
    const auto &e = a.e;
    const int B = a.batch, L = a.limbs;
    scale_ok(*e, a.scale * b_scale, L);
    Batch prod(e, B, 3, L, a.scale * b_scale), lin(e, B, 2, L, prod.scale);
    ckks_view va = a.view(), vp = prod.view(), vl = lin.view();
    check(ckks_multiply(e->ctx, &va, &vb, &vp, nullptr));
    check(ckks_relinearize(e->ctx, &vp, rk.s->keys.at(0)->p, &vl, nullptr));
    Batch mult(e, B, 2, L - 1, prod.scale / (double)e->primes[L - 1]);
    ckks_view vm = mult.view();
    check(ckks_rescale(e->ctx, &vl, &vm, nullptr));
    Batch rot(e, B, 2, L - 1, mult.scale), scratch(e, B, 2, L - 1, mult.scale), dup(e, B, 2, L - 1, mult.scale);
    ckks_view vr = rot.view(), vs = scratch.view(), vd = dup.view();
    check(ckks_rotate(e->ctx, gk.s->ks, &vm, -size, &vr, &vs, nullptr));          // "vector has zeros now"
    check(ckks_add(e->ctx, &vm, &vr, &vd, nullptr));                               // "vector has duplicate now"
    if (size > 1) {
        int final_in_b = 0;
        check(ckks_rotate_sum_chain(e->ctx, gk.s->ks, &vd, &vs, &vm, 1, size - 1, &final_in_b, nullptr));
    }
    mult.scale = std::pow(2.0, (int)std::log2(mult.scale));
    return mult;
}

inline Batch dot_product_batch(const Batch &a, const Poly &b, int size, const seal::RelinKeys &rk, const seal::GaloisKeys &gk) {
    if (!b.buf) throw std::invalid_argument("encrypted is not valid for encryption parameters");
    return dot_product_batch(a, b.view(), b.scale, size, rk, gk);
}

// one-hot masks e_0..e_{count-1} (length `length`), encoded as one batch at `limbs`
inline Batch one_hot_masks(const std::shared_ptr<Engine> &e, int count, int length, double scale, int limbs) {
    std::vector<double> m((std::size_t)count * length, 0.0);
    for (int i = 0; i < count; i++) m[(std::size_t)i * length + i] = 1.0;
    DevBuf vals(e, m.size());
    check(ckks_upload(e->ctx, vals.p, m.data(), m.size() * 8, nullptr));
    check(ckks_stream_sync(e->ctx, nullptr));
    Batch pts(e, count, 1, limbs, scale);
    ckks_view vp = pts.view();
    check(ckks_encode(e->ctx, reinterpret_cast<const double *>(vals.p), length, scale, &vp, nullptr));
    return pts;
}

// multiply_plain_inplace of every entry with its mask, add_many, rescale, "manual rescale" of the scale
inline seal::Ciphertext masked_sum(const Batch &cts, const Batch &masks, seal::Evaluator &evaluator) {
    scale_ok(*cts.e, cts.scale * masks.scale, cts.limbs);
    seal::Ciphertext sum;
    sum.poly().allocate(cts.e, 2, cts.limbs);
    sum.poly().scale = cts.scale * masks.scale;
    ckks_view vc = cts.view(), vm = masks.view(), vo = sum.poly().view();
    check(ckks_multiply_plain_sum(cts.e->ctx, &vc, &vm, &vo, nullptr));
    evaluator.rescale_to_next_inplace(sum);          // relinearize_inplace on a size-2 ciphertext is a no-op (:236, :318)
    sum.scale() = std::pow(2.0, (int)std::log2(sum.scale()));
    return sum;
}

}  // namespace detail

// logistic_regression_ckks.cpp:208-266 -- sigmoid polynomial of the R per-row dot products, slot i = row i
inline seal::Ciphertext predict_cipher_weights(const std::vector<seal::Ciphertext> &features, const seal::Ciphertext &weights,
                                               int num_weights, double scale, seal::Evaluator &evaluator,
                                               seal::CKKSEncoder &ckks_encoder, const seal::GaloisKeys &gal_keys,
                                               const seal::RelinKeys &relin_keys, seal::Encryptor &encryptor,
                                               const seal::EncryptionParameters &params, int degree = 3, bool tree = false) {
    const int num_rows = (int)features.size();
    detail::Batch rows = detail::gather(features);
    if (!weights.poly().buf) throw std::invalid_argument("encrypted is not valid for encryption parameters");
    detail::Batch results = detail::dot_product_batch(rows, weights.poly(), num_weights, relin_keys, gal_keys);
    // masks are encoded at the top level and mod-switched to the next one (:225-227) = encoded at that level
    detail::Batch masks = detail::one_hot_masks(rows.e, num_rows, num_rows, scale, rows.limbs - 1);
    seal::Ciphertext lintransf_vec = detail::masked_sum(results, masks, evaluator);
    std::vector<double> coeffs = sigmoid_coeffs(degree);
    if (tree) return Tree_cipher(lintransf_vec, degree, scale, coeffs, ckks_encoder, evaluator, encryptor, relin_keys, params);
    return Horner_cipher(lintransf_vec, degree, coeffs, ckks_encoder, scale, evaluator, encryptor, relin_keys, params);
}

// the tail of update_weights (logistic_regression_ckks.cpp:326-342, repair R3): weights - learning_rate / R * gradient
inline seal::Ciphertext apply_gradient(seal::Ciphertext gradient, const seal::Ciphertext &weights, double learning_rate,
                                       int num_observations, double scale, seal::Evaluator &evaluator, seal::CKKSEncoder &ckks_encoder) {
    seal::Plaintext N_pt;
    ckks_encoder.encode(learning_rate / num_observations, scale, N_pt);                          // :330-333
    evaluator.mod_switch_to_inplace(N_pt, gradient.parms_id());
    evaluator.multiply_plain_inplace(gradient, N_pt);
    evaluator.rescale_to_next_inplace(gradient);                                                 // R3
    gradient.scale() = std::pow(2.0, (int)std::log2(gradient.scale()));
    seal::Ciphertext w_low = weights;
    evaluator.mod_switch_to_inplace(w_low, gradient.parms_id());                                 // R3
    w_low.scale() = gradient.scale();
    seal::Ciphertext new_weights;
    evaluator.sub(gradient, w_low, new_weights);                                                 // :341
    evaluator.negate_inplace(new_weights);                                                       // :342
    return new_weights;
}

// logistic_regression_ckks.cpp:269-345 -- weights - learning_rate / R * X^T (sigmoid(X w) - y), with repair R3
inline seal::Ciphertext update_weights(const std::vector<seal::Ciphertext> &features, const std::vector<seal::Ciphertext> &features_T,
                                       seal::Ciphertext labels, const seal::Ciphertext &weights, float learning_rate,
                                       seal::Evaluator &evaluator, seal::CKKSEncoder &ckks_encoder, const seal::GaloisKeys &gal_keys,
                                       const seal::RelinKeys &relin_keys, seal::Encryptor &encryptor, double scale,
                                       const seal::EncryptionParameters &params, int degree = 3, bool tree = false) {
    const int num_observations = (int)features.size(), num_weights = (int)features_T.size();
    seal::Ciphertext predictions = predict_cipher_weights(features, weights, num_weights, scale, evaluator, ckks_encoder, gal_keys,
                                                          relin_keys, encryptor, params, degree, tree);
    evaluator.mod_switch_to_inplace(labels, predictions.parms_id());
    predictions.scale() = labels.scale();             // both are 2^40 after the forced scales
    seal::Ciphertext pred_labels;
    evaluator.sub(predictions, labels, pred_labels);
    const int Lg = (int)pred_labels.coeff_mod_count();
    detail::Batch cols = detail::gather_at(features_T, Lg);                                      // :295
    detail::Batch grads = detail::dot_product_batch(cols, pred_labels.poly(), num_observations, relin_keys, gal_keys);
    detail::Batch masks = detail::one_hot_masks(cols.e, num_weights, num_weights, scale, grads.limbs);   // :305-308
    seal::Ciphertext gradient = detail::masked_sum(grads, masks, evaluator);
    return apply_gradient(gradient, weights, learning_rate, num_observations, scale, evaluator, ckks_encoder);
}

// ---- config 5 (BASELINE.json): 8 features x 32768 samples do not fit the one-ciphertext-per-row layout (the mask
// would be longer than the slot count), so samples are held as mini-batches of B <= slots/2 samples, each C column
// ciphertexts; the weights are supplied as C broadcast ciphertexts.  Mini-batches are the multi-GPU sharding unit.
struct ColumnLayout {
    int R, C, B, M;
    std::size_t slots;
    ColumnLayout(int rows, int cols, int batch, std::size_t slot_count) : R(rows), C(cols), B(batch), M(rows / batch), slots(slot_count) {
        if (R % B || (std::size_t)(2 * B) > slots) throw std::invalid_argument("B must divide R and 2B must fit in N/2 slots");
    }
    // entry m*C + j = feature j of mini-batch m
    std::vector<double> column(const std::vector<std::vector<double>> &X, int m, int j) const {
        std::vector<double> v(slots, 0.0);
        for (int i = 0; i < B; i++) v[i] = X[(std::size_t)m * B + i][j];
        return v;
    }
    std::vector<double> labels(const std::vector<double> &y, int m) const {
        std::vector<double> v(slots, 0.0);
        for (int i = 0; i < B; i++) v[i] = y[(std::size_t)m * B + i];
        return v;
    }
};

// One pass of the gradient loop over M mini-batches (the evaluator sequence of lr.py:column_epoch_gradient):
// per mini-batch z = sum_j multiply(col_j, w_j), relinearize, rescale, sigmoid polynomial, sub labels; then, per
// (mini-batch, feature), the reference's cipher_dot_product over B slots -- all M*C chains in lock-step -- the one-hot
// mask e_j, add_many over everything, rescale.  Slot j of the result = sum_i x_ij (sigmoid(x_i . w) - y_i).
inline seal::Ciphertext column_epoch_gradient(const std::vector<seal::Ciphertext> &cols, const std::vector<seal::Ciphertext> &labels,
                                              const std::vector<seal::Ciphertext> &w_bcast, int B, double scale,
                                              seal::Evaluator &evaluator, seal::CKKSEncoder &ckks_encoder,
                                              const seal::GaloisKeys &gal_keys, const seal::RelinKeys &relin_keys,
                                              seal::Encryptor &encryptor, const seal::EncryptionParameters &params, int degree = 7,
                                              bool tree = true) {
    const int C = (int)w_bcast.size(), M = (int)labels.size();
    if (C < 1 || M < 1 || (int)cols.size() != M * C) throw std::invalid_argument("encrypted1 and encrypted2 parameter mismatch");
    detail::Batch colb = detail::gather(cols), wb = detail::gather(w_bcast);
    const auto &e = colb.e;
    const int L = colb.limbs;
    if (wb.limbs != L) throw std::invalid_argument("encrypted1 and encrypted2 parameter mismatch");
    if (!relin_keys.s || !relin_keys.s->keys.count(0)) throw std::invalid_argument("relin_keys is not valid for encryption parameters");
    detail::scale_ok(*e, colb.scale * wb.scale, L);
    // z_m = sum_j col_{m,j} x w_j: the fused multiply + add_many over the C features of each mini-batch
    detail::Batch z3(e, M, 3, L, colb.scale * wb.scale), z2(e, M, 2, L, z3.scale), z(e, M, 2, L - 1, z3.scale / (double)e->primes[L - 1]);
    for (int m = 0; m < M; m++) {
        ckks_view va = colb.view(), vw = wb.view(), vo = z3.view();
        va.data += (std::size_t)m * C * colb.entry_words();
        va.batch = C;
        vo.data += (std::size_t)m * z3.entry_words();
        vo.batch = 1;
        detail::check(ckks_multiply_sum(e->ctx, &va, &vw, &vo, nullptr));
    }
    ckks_view v3 = z3.view(), v2 = z2.view(), vz = z.view();
    detail::check(ckks_relinearize(e->ctx, &v3, relin_keys.s->keys.at(0)->p, &v2, nullptr));
    detail::check(ckks_rescale(e->ctx, &v2, &vz, nullptr));
    z.scale = std::pow(2.0, (int)std::log2(z.scale));
    // sigmoid polynomial and label subtraction per mini-batch (a handful of sequential ops each)
    std::vector<double> coeffs = sigmoid_coeffs(degree);
    std::vector<seal::Ciphertext> pred_labels(M);
    for (int m = 0; m < M; m++) {
        seal::Ciphertext zm;
        z.get(m, zm.poly());
        seal::Ciphertext pred = tree ? Tree_cipher(zm, degree, scale, coeffs, ckks_encoder, evaluator, encryptor, relin_keys, params)
                                     : Horner_cipher(zm, degree, coeffs, ckks_encoder, scale, evaluator, encryptor, relin_keys, params);
        seal::Ciphertext lab = labels[m];
        evaluator.mod_switch_to_inplace(lab, pred.parms_id());
        pred.scale() = lab.scale();
        evaluator.sub(pred, lab, pred_labels[m]);
    }
    const int Lg = (int)pred_labels[0].coeff_mod_count();
    detail::Batch colv = detail::gather_at(cols, Lg);
    detail::Batch pl(e, M * C, 2, Lg, pred_labels[0].scale());              // entry m*C + j = pred_labels[m]
    for (int m = 0; m < M; m++)
        for (int j = 0; j < C; j++) pl.put(m * C + j, pred_labels[m].poly());
    detail::Batch grads = detail::dot_product_batch(colv, pl.view(), pl.scale, B, relin_keys, gal_keys);
    // masks e_j for every (m, j)
    std::vector<double> mv((std::size_t)M * C * C, 0.0);
    for (int m = 0; m < M; m++)
        for (int j = 0; j < C; j++) mv[((std::size_t)m * C + j) * C + j] = 1.0;
    detail::DevBuf vals(e, mv.size());
    detail::check(ckks_upload(e->ctx, vals.p, mv.data(), mv.size() * 8, nullptr));
    detail::check(ckks_stream_sync(e->ctx, nullptr));
    detail::Batch masks(e, M * C, 1, grads.limbs, scale);
    ckks_view vm = masks.view();
    detail::check(ckks_encode(e->ctx, reinterpret_cast<const double *>(vals.p), C, scale, &vm, nullptr));
    return detail::masked_sum(grads, masks, evaluator);
}

// logistic_regression_ckks.cpp:348-385 -- `iters` training iterations; the weights are refreshed after each one
// by the party holding the secret key (decrypt, decode, re-encode, encrypt).  Repair R4: the reference
// re-encrypts the low-level plaintext it just decrypted, which leaves no levels for the next iteration; here
// the decoded weights are re-packed periodically (RowLayout::weights) and encoded at the top level.
inline seal::Ciphertext train_cipher(const std::vector<seal::Ciphertext> &features, const std::vector<seal::Ciphertext> &features_T,
                                     const seal::Ciphertext &labels, const seal::Ciphertext &weights, float learning_rate, int iters,
                                     int observations, int num_weights, seal::Evaluator &evaluator, seal::CKKSEncoder &ckks_encoder,
                                     double scale, const seal::GaloisKeys &gal_keys, const seal::RelinKeys &relin_keys,
                                     seal::Encryptor &encryptor, seal::Decryptor &decryptor, const seal::EncryptionParameters &params,
                                     int degree = 3, bool tree = false) {
    RowLayout layout(observations, num_weights, ckks_encoder.slot_count());
    seal::Ciphertext new_weights = weights;
    for (int i = 0; i < iters; i++) {
        new_weights = update_weights(features, features_T, labels, new_weights, learning_rate, evaluator, ckks_encoder, gal_keys,
                                     relin_keys, encryptor, scale, params, degree, tree);
        seal::Plaintext new_weights_pt;
        decryptor.decrypt(new_weights, new_weights_pt);
        std::vector<double> decoded;
        ckks_encoder.decode(new_weights_pt, decoded);
        decoded.resize(num_weights);
        seal::Plaintext fresh;
        ckks_encoder.encode(layout.weights(decoded), scale, fresh);
        encryptor.encrypt(fresh, new_weights);
    }
    return new_weights;
}

}  // namespace b200
