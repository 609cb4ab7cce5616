"""Sequential restatement of the reference's layer-2 algorithms on top of the CPU oracle, written
the way the reference writes them (one evaluator call per line, no batching, no sharing), so the
batched GPU workloads can be compared bit-for-bit.  Test infrastructure only."""
import math

import numpy as np


class OCt:
    """oracle-side ciphertext / plaintext: numpy limbs + scale"""

    def __init__(self, data, scale):
        self.data, self.scale = data, float(scale)

    @property
    def limbs(self):
        return self.data.shape[-2]

    def copy(self):
        return OCt(self.data.copy(), self.scale)


class OEval:
    """SEAL-like evaluator over the oracle"""

    def __init__(self, orc, rlk, gks):
        self.o, self.rlk, self.gks = orc, rlk, gks

    def add(self, a, b):
        assert a.scale == b.scale and a.limbs == b.limbs
        return OCt(self.o.add(a.data, b.data), a.scale)

    def sub(self, a, b):
        assert a.scale == b.scale and a.limbs == b.limbs
        return OCt(self.o.sub(a.data, b.data), a.scale)

    def negate(self, a):
        return OCt(self.o.negate(a.data), a.scale)

    def add_many(self, cts):
        acc = cts[0]
        for c in cts[1:]:
            acc = self.add(acc, c)
        return acc

    def multiply(self, a, b):
        return OCt(self.o.multiply(a.data, b.data), a.scale * b.scale)

    def multiply_plain(self, a, p):
        return OCt(self.o.multiply_plain(a.data, p.data), a.scale * p.scale)

    def add_plain(self, a, p):
        assert a.scale == p.scale
        return OCt(self.o.add_plain(a.data, p.data), a.scale)

    def relinearize(self, a):
        if a.data.shape[0] == 2:
            return a
        return OCt(self.o.relinearize(a.data, self.rlk), a.scale)

    def rescale(self, a):
        return OCt(self.o.rescale(a.data), a.scale / self.o.primes[a.limbs - 1])

    def mod_switch_to(self, a, limbs):
        assert limbs <= a.limbs
        return OCt(np.ascontiguousarray(a.data[..., :limbs, :]), a.scale)

    def rotate(self, a, steps):
        return OCt(self.o.rotate(a.data, steps, self.gks), a.scale)


def pow2(x):
    return float(2.0 ** int(math.log2(x)))


def linear_transform_plain(E, ct, diags):
    """helper.h:237-262"""
    d = len(diags)
    ct_new = E.add(ct, E.rotate(ct, -d))
    res = [E.multiply_plain(ct_new, diags[0])]
    for l in range(1, d):
        res.append(E.multiply_plain(E.rotate(ct_new, l), diags[l]))
    return E.add_many(res)


def linear_transform_cipher(E, ct, diag_cts):
    """helper.h:212-234"""
    d = len(diag_cts)
    ct_new = E.add(ct, E.rotate(ct, -d))
    res = [E.multiply(ct_new, diag_cts[0])]
    for l in range(1, d):
        res.append(E.multiply(E.rotate(ct_new, l), diag_cts[l]))
    return E.add_many(res)


def c_matrix_encode(E, rows):
    """helper.h:307-322"""
    d = len(rows)
    rots = [rows[0]] + [E.rotate(rows[i], -(i * d)) for i in range(1, d)]
    return E.add_many(rots)


def cc_matrix_multiplication(E, ctA, ctB, d, sig, tau, V, W):
    """matrix_mult_benchmark.cpp:13-71"""
    A = [linear_transform_plain(E, ctA, sig)]
    B = [linear_transform_plain(E, ctB, tau)]
    for k in range(1, d):
        A.append(linear_transform_plain(E, A[0], V[k - 1]))
        B.append(linear_transform_plain(E, B[0], W[k - 1]))
    for i in range(1, d):
        A[i] = E.rescale(A[i])
        B[i] = E.rescale(B[i])
    ctAB = E.multiply(A[0], B[0])
    ctAB = E.mod_switch_to(ctAB, ctAB.limbs - 1)
    for i in range(1, d):
        A[i].scale = pow2(A[i].scale)
        B[i].scale = pow2(B[i].scale)
    for k in range(1, d):
        ctAB = E.add(ctAB, E.multiply(A[k], B[k]))
    return ctAB


def cipher_dot_product(E, a, b, size):
    """helper.h:416-502"""
    mult = E.rescale(E.relinearize(E.multiply(a, b)))
    dup = E.add(mult, E.rotate(mult, -size))
    for _ in range(1, size):
        dup = E.rotate(dup, 1)
        mult = E.add(mult, dup)
    mult.scale = pow2(mult.scale)
    return mult


def compute_all_powers(E, x, degree):
    """helper.h:505-547"""
    powers = [None] * (degree + 1)
    powers[1] = x
    levels = [0] * (degree + 1)
    for i in range(2, degree + 1):
        minlevel, cand = i, -1
        for j in range(1, i // 2 + 1):
            nl = max(levels[j], levels[i - j]) + 1
            if nl < minlevel:
                cand, minlevel = j, nl
        levels[i] = minlevel
        temp = E.mod_switch_to(powers[cand], powers[i - cand].limbs)
        powers[i] = E.rescale(E.relinearize(E.multiply(temp, powers[i - cand])))
    return powers


def horner_cipher(E, x, coeffs, scale, encode, encrypt):
    """logistic_regression_ckks.cpp:139-205"""
    degree = len(coeffs) - 1
    temp = encrypt(encode(float(coeffs[degree]), scale, None))
    for i in range(degree - 1, -1, -1):
        if x.limbs > temp.limbs:
            x = E.mod_switch_to(x, temp.limbs)
        elif x.limbs < temp.limbs:
            temp = E.mod_switch_to(temp, x.limbs)
        temp = E.rescale(E.relinearize(E.multiply(temp, x)))
        temp.scale = float(2.0 ** 40)
        temp = E.add_plain(temp, encode(float(coeffs[i]), scale, temp.limbs))
    return temp


def tree_cipher(E, x, coeffs, scale, encode, encrypt):
    """logistic_regression_ckks.cpp:55-137"""
    degree = len(coeffs) - 1
    powers = compute_all_powers(E, x, degree)
    res = encrypt(encode(float(coeffs[0]), scale, None))
    for i in range(1, degree + 1):
        temp = E.rescale(E.multiply_plain(powers[i], encode(float(coeffs[i]), scale, powers[i].limbs)))
        res = E.mod_switch_to(res, temp.limbs)
        res.scale = pow2(res.scale)
        temp.scale = pow2(res.scale)
        res = E.add(res, temp)
    return res


def update_weights(E, features, features_T, labels, weights, lr, scale, coeffs, encode, encrypt, method):
    """logistic_regression_ckks.cpp:208-345 with the repairs listed in the product's lr.py"""
    R, C = len(features), len(features_T)
    results = []
    for i in range(R):
        r = cipher_dot_product(E, features[i], weights, C)
        mask = np.zeros(R)
        mask[i] = 1.0
        mpt = encode(mask, scale, None)
        mpt = OCt(mpt.data[:-1], mpt.scale)                 # mod_switch_to_next_inplace(mask_pt)
        results.append(E.multiply_plain(r, mpt))
    lin = E.rescale(E.relinearize(E.add_many(results)))
    lin.scale = pow2(lin.scale)
    poly = tree_cipher if method == "tree" else horner_cipher
    pred = poly(E, lin, coeffs, scale, encode, encrypt)
    lab = E.mod_switch_to(labels, pred.limbs)
    pred.scale = lab.scale
    pred_labels = E.sub(pred, lab)
    grads = []
    for j in range(C):
        col = E.mod_switch_to(features_T[j], pred_labels.limbs)
        g = cipher_dot_product(E, col, pred_labels, R)
        mask = np.zeros(C)
        mask[j] = 1.0
        grads.append(E.multiply_plain(g, encode(mask, scale, g.limbs)))
    gradient = E.rescale(E.relinearize(E.add_many(grads)))
    gradient.scale = pow2(gradient.scale)
    gradient = E.multiply_plain(gradient, encode(float(lr / R), scale, gradient.limbs))
    gradient = E.rescale(gradient)
    gradient.scale = pow2(gradient.scale)
    w_low = E.mod_switch_to(weights, gradient.limbs)
    w_low.scale = gradient.scale
    return E.negate(E.sub(gradient, w_low))


def linear_transform_ciphermatrix_plainvector(E, pt_rotations, ct_diags):
    """helper.h:265-278: sum_i multiply_plain(ct_diag_i, pt_rot_i)"""
    return E.add_many([E.multiply_plain(c, p) for c, p in zip(ct_diags, pt_rotations)])


def c_matrix_decode(E, matrix, d, scale, encode):
    """helper.h:325-360: row i = rotate(multiply_plain(matrix, ones on [i*d, (i+1)*d)), i*d)"""
    rows = []
    for i in range(d):
        mask = np.zeros(d * d)
        mask[i * d:(i + 1) * d] = 1.0
        row = E.multiply_plain(matrix, encode(mask, scale, matrix.limbs))
        rows.append(row if i == 0 else E.rotate(row, i * d))
    return rows


def column_epoch_gradient(E, cols, labels, w_bcast, C, B, scale, coeffs, encode, encrypt_for_batch, method="tree"):
    """Sequential restatement of the product's column-layout epoch (lr.py column_epoch_gradient): per
    mini-batch m the prediction z = sum_j multiply(col_j, w_j), relinearize, rescale, forced scale, the
    sigmoid polynomial (logistic_regression_ckks.cpp:55-137 / :139-205), sub labels (:286-288); per
    feature j the reference's cipher_dot_product over B slots and the one-hot mask e_j (:295-311);
    add_many over every (m, j); rescale; forced scale (:316-323).
    cols: list of M*C OCt (entry m*C + j); labels: list of M; w_bcast: list of C.
    encrypt_for_batch(m) returns the encrypt callable of mini-batch m (the batched product path draws
    one fresh encryption per polynomial call and shares it across the batch)."""
    M = len(labels)
    poly = tree_cipher if method == "tree" else horner_cipher
    grads = []
    for m in range(M):
        prods = [E.multiply(cols[m * C + j], w_bcast[j]) for j in range(C)]
        z = E.rescale(E.relinearize(E.add_many(prods)))
        z.scale = pow2(z.scale)
        pred = poly(E, z, coeffs, scale, encode, encrypt_for_batch(m))
        lab = E.mod_switch_to(labels[m], pred.limbs)
        pred.scale = lab.scale
        pred_labels = E.sub(pred, lab)
        for j in range(C):
            col = E.mod_switch_to(cols[m * C + j], pred_labels.limbs)
            g = cipher_dot_product(E, col, pred_labels, B)
            mask = np.zeros(C)
            mask[j] = 1.0
            grads.append(E.multiply_plain(g, encode(mask, scale, g.limbs)))
    gradient = E.rescale(E.add_many(grads))
    gradient.scale = pow2(gradient.scale)
    return gradient


def apply_gradient(E, gradient, weights, lr, R, scale, encode):
    """tail of update_weights (logistic_regression_ckks.cpp:326-342, repair R3)"""
    g = E.rescale(E.multiply_plain(gradient, encode(float(lr / R), scale, gradient.limbs)))
    g.scale = pow2(g.scale)
    w_low = E.mod_switch_to(weights, g.limbs)
    w_low.scale = g.scale
    return E.negate(E.sub(g, w_low))
