"""GPU parity of the batched workloads (Linear_Transform_*, C_Matrix_Encode, CC_Matrix_Multiplication,
cipher_dot_product, Horner/Tree polynomial, LR update) against the sequential oracle restatement
of the reference's loops, bit-exact on the final ciphertext; plus decrypt-and-compare against
plaintext math within the stated tolerance."""
import importlib

import numpy as np
import pytest

import ref_workloads as rw

pytestmark = pytest.mark.gpu
PKG = "seal-fyp-logistic-regression_b200"


class Bridge:
    """oracle-backed encoder / encryptor adapters so that the GPU workloads and the oracle sequence
    consume identical plaintexts and identical fresh encryptions"""

    def __init__(self, fx):
        self.fx = fx
        self.seed = 1000

    # product-side interfaces ------------------------------------------------
    def encode(self, values, scale, limbs=None):
        L = self.fx.L if limbs is None else limbs
        if np.isscalar(values):
            return self.fx.ctx.upload_plain(self.fx.orc.encode(values, scale, L), scale=scale)
        vals = np.atleast_2d(values)
        arr = np.stack([self.fx.orc.encode(v, scale, L) for v in vals])
        return self.fx.ctx.upload_plain(arr, scale=scale)

    def encrypt(self, pt):
        self.seed += 1
        arr = pt.numpy()[:, 0]
        cts = np.stack([self.fx.orc.encrypt(self.seed, self.fx.pk, a) for a in arr])
        return self.fx.ctx.upload(cts, scale=pt.scale)

    # oracle-side twins (same seeds in the same order) --------------------------
    def o_encode(self, values, scale, limbs):
        L = self.fx.L if limbs is None else limbs
        return rw.OCt(self.fx.orc.encode(values, scale, L), scale)

    def o_encrypt(self, pt):
        self.seed += 1
        return rw.OCt(self.fx.orc.encrypt(self.seed, self.fx.pk, pt.data), pt.scale)


def _mods():
    return (importlib.import_module(PKG + ".workloads"), importlib.import_module(PKG + ".lr"),
            importlib.import_module(PKG + ".client"))


def _enc(fx, seed, values, scale, L=None):
    pt = fx.orc.encode(values, scale, fx.L if L is None else L)
    return fx.orc.encrypt(seed, fx.pk, pt)


POW2 = tuple(s for i in range(11) for s in (1 << i, -(1 << i)))


@pytest.fixture(scope="module")
def fx12(make_fixture):
    return make_fixture(12, [50, 40, 40, 50], steps=POW2)


def test_linear_transform_plain_and_cipher(fx12):
    wl, _, _ = _mods()
    fx = fx12
    E = rw.OEval(fx.orc, fx.rlk, fx.gks)
    plans = wl.PlanCache(fx.ctx, fx.keys)
    rng = np.random.default_rng(1)
    scale = 2.0 ** 40
    for d in (5, 8, 13):
        U = rng.uniform(0, 1, (d, d))
        v = rng.uniform(0, 1, d)
        diags = wl.all_diagonals(U)
        ct = _enc(fx, 50 + d, v, scale)
        pts = np.stack([fx.orc.encode(dg, scale) for dg in diags])
        want = rw.linear_transform_plain(E, rw.OCt(ct, scale), [rw.OCt(p, scale) for p in pts])
        got = wl.linear_transform_plain(fx.ev, fx.ctx.upload(ct, scale=scale), fx.ctx.upload_plain(pts, scale=scale),
                                        fx.keys, plans)
        assert got.scale == want.scale
        assert np.array_equal(got.numpy()[0], want.data), d
        dec = fx.orc.decode(fx.orc.decrypt(fx.sk, got.numpy()[0]), got.scale)[:d]
        assert np.abs(dec - U @ v).max() < 1e-4          # test_Linear_Transformation (linear_transformation.cpp:203-218)
    # ciphertext diagonals (Linear_Transform_Cipher)
    d = 6
    U = rng.uniform(0, 1, (d, d))
    v = rng.uniform(0, 1, d)
    diags = wl.all_diagonals(U)
    ct = _enc(fx, 70, v, scale)
    dcts = np.stack([_enc(fx, 71 + i, diags[i], scale) for i in range(d)])
    want = rw.linear_transform_cipher(E, rw.OCt(ct, scale), [rw.OCt(c, scale) for c in dcts])
    got = wl.linear_transform_cipher(fx.ev, fx.ctx.upload(ct, scale=scale), fx.ctx.upload(dcts, scale=scale), fx.keys, plans)
    assert got.size == 3 and np.array_equal(got.numpy()[0], want.data)
    dec = fx.orc.decode(fx.orc.decrypt(fx.sk, got.numpy()[0]), got.scale)[:d]
    assert np.abs(dec - U @ v).max() < 1e-4


def test_linear_transform_bsgs_mode(fx12):
    """baby-step / giant-step evaluation of Linear_Transform_Plain -- different ciphertext
    polynomials, same decrypted vector as the reference sequence (and as U @ v), far fewer key switches"""
    wl, _, client = _mods()
    fx = fx12
    plans = wl.PlanCache(fx.ctx, fx.keys)
    enc = client.CKKSEncoder(fx.ctx)
    rng = np.random.default_rng(12)
    scale = 2.0 ** 40
    for d, baby in ((5, None), (13, 4), (32, None), (64, 8)):
        U, v = rng.uniform(0, 1, (d, d)), rng.uniform(0, 1, d)
        ct = fx.ctx.upload(_enc(fx, 300 + d, v, scale), scale=scale)
        ref = wl.linear_transform_plain(fx.ev, ct, enc.encode(wl.all_diagonals(U), scale), fx.keys, plans)
        bd = wl.BsgsDiagonals(U, scale, enc, baby=baby)
        got = wl.linear_transform_plain_bsgs(fx.ev, ct, bd, fx.keys, plans)
        assert got.limbs == ref.limbs and got.scale == ref.scale
        assert not np.array_equal(got.numpy(), ref.numpy())
        dg = fx.orc.decode(fx.orc.decrypt(fx.sk, got.numpy()[0]), got.scale)[:d]
        dr = fx.orc.decode(fx.orc.decrypt(fx.sk, ref.numpy()[0]), ref.scale)[:d]
        assert np.abs(dg - dr).max() < 1e-5 and np.abs(dg - U @ v).max() < 1e-4, d
        ks_ref = plans.get(range(d)).keyswitches
        ks_bsgs = plans.get(range(bd.b)).keyswitches + plans.get([g * bd.b for g in range(bd.G)]).keyswitches
        assert ks_bsgs < ks_ref or d <= 5


def test_matrix_encode_and_multiplication(make_fixture):
    wl, _, _ = _mods()
    fx = make_fixture(12, [50, 40, 40, 40, 40, 50], steps=POW2)
    E = rw.OEval(fx.orc, fx.rlk, fx.gks)
    plans = wl.PlanCache(fx.ctx, fx.keys)
    rng = np.random.default_rng(2)
    scale = 2.0 ** 40
    d = 3
    dd = d * d
    A, B = rng.uniform(0, 1, (d, d)), rng.uniform(0, 1, (d, d))
    eps = 1e-8                                        # matrix_multiplication.cpp:239-246

    def enc_diags(U):
        return np.stack([fx.orc.encode(dg + eps, scale) for dg in wl.all_diagonals(U)])

    sig, tau = enc_diags(wl.u_sigma(d)), enc_diags(wl.u_tau(d))
    V = [enc_diags(wl.v_k(d, k)) for k in range(1, d)]
    W = [enc_diags(wl.w_k(d, k)) for k in range(1, d)]
    rowsA = np.stack([_enc(fx, 80 + i, A[i], scale) for i in range(d)])
    rowsB = np.stack([_enc(fx, 90 + i, B[i], scale) for i in range(d)])
    # C_Matrix_Encode
    wantA = rw.c_matrix_encode(E, [rw.OCt(r, scale) for r in rowsA])
    wantB = rw.c_matrix_encode(E, [rw.OCt(r, scale) for r in rowsB])
    gotA = wl.c_matrix_encode(fx.ev, fx.ctx.upload(rowsA, scale=scale), fx.keys, plans)
    gotB = wl.c_matrix_encode(fx.ev, fx.ctx.upload(rowsB, scale=scale), fx.keys, plans)
    assert np.array_equal(gotA.numpy()[0], wantA.data) and np.array_equal(gotB.numpy()[0], wantB.data)
    oc = lambda arrs: [rw.OCt(a, scale) for a in arrs]
    want = rw.cc_matrix_multiplication(E, wantA, wantB, d, oc(sig), oc(tau), [oc(v) for v in V], [oc(w) for w in W])
    up = lambda arrs: fx.ctx.upload_plain(arrs, scale=scale)
    got = wl.cc_matrix_multiplication(fx.ev, gotA, gotB, d, up(sig), up(tau), [up(v) for v in V], [up(w) for w in W],
                                      fx.keys, plans)
    assert got.limbs == want.limbs and got.scale == want.scale
    assert np.array_equal(got.numpy()[0], want.data)
    dec = fx.orc.decode(fx.orc.decrypt(fx.sk, got.numpy()[0]), got.scale)[:dd].reshape(d, d)
    assert np.abs(dec - A @ B).max() < 1e-3           # test_matrix_mult (matrix_mult_benchmark.cpp:73-88)
    # permutation matrices: closed forms act as the paper says
    a = np.arange(dd, dtype=float).reshape(d, d)
    assert np.array_equal((wl.u_transpose(d) @ a.reshape(-1)).reshape(d, d), a.T)
    assert np.array_equal((wl.v_k(d, 1) @ a.reshape(-1)).reshape(d, d), np.roll(a, -1, axis=1))
    assert np.array_equal((wl.w_k(d, 1) @ a.reshape(-1)).reshape(d, d), np.roll(a, -1, axis=0))


def test_dot_product_and_polynomials(make_fixture):
    wl, lr, _ = _mods()
    fx = make_fixture(12, [50, 40, 40, 40, 40, 50], steps=POW2)
    E = rw.OEval(fx.orc, fx.rlk, fx.gks)
    br = Bridge(fx)
    rng = np.random.default_rng(3)
    scale = 2.0 ** 40
    size = 8
    a, b = rng.uniform(-1, 1, (3, size)), rng.uniform(-1, 1, (3, size))
    ca = np.stack([_enc(fx, 100 + i, a[i], scale) for i in range(3)])
    cb = np.stack([_enc(fx, 110 + i, b[i], scale) for i in range(3)])
    got = wl.cipher_dot_product(fx.ev, fx.ctx.upload(ca, scale=scale), fx.ctx.upload(cb, scale=scale), size, fx.keys)
    for i in range(3):
        want = rw.cipher_dot_product(E, rw.OCt(ca[i], scale), rw.OCt(cb[i], scale), size)
        assert np.array_equal(got.numpy()[i], want.data) and got.scale == want.scale
        dec = fx.orc.decode(fx.orc.decrypt(fx.sk, got.numpy()[i]), got.scale)[:size]
        assert np.abs(dec - a[i] @ b[i]).max() < 1e-4
    # polynomials: Horner degree 3 and tree degree 7 with the reference's sigmoid coefficients
    x = rng.uniform(-1, 1, 32)
    cx = _enc(fx, 120, x, scale)
    for method, degree in (("horner", 3), ("tree", 7)):
        coeffs = lr.SIGMOID_COEFFS[degree]
        br.seed = 2000
        fn = wl.tree_cipher if method == "tree" else wl.horner_cipher
        got = fn(fx.ev, fx.ctx.upload(cx, scale=scale), coeffs, scale, fx.keys, br, br)
        br.seed = 2000
        ofn = rw.tree_cipher if method == "tree" else rw.horner_cipher
        want = ofn(E, rw.OCt(cx, scale), coeffs, scale, br.o_encode, br.o_encrypt)
        assert got.limbs == want.limbs and got.scale == want.scale
        assert np.array_equal(got.numpy()[0], want.data), method
        dec = fx.orc.decode(fx.orc.decrypt(fx.sk, got.numpy()[0]), got.scale)[:32]
        expect = sum(c * x ** i for i, c in enumerate(coeffs))
        assert np.abs(dec - expect).max() < 1e-3, method


def test_lr_update_weights_row_layout(make_fixture):
    """one repaired update_weights (logistic_regression_ckks.cpp:269-345), R = 12 rows x C = 4
    features, Horner degree 3: bit-exact vs the sequential oracle and close to plaintext LR"""
    wl, lr, _ = _mods()
    fx = make_fixture(12, [50] + [40] * 8 + [50], steps=POW2)
    E = rw.OEval(fx.orc, fx.rlk, fx.gks)
    br = Bridge(fx)
    rng = np.random.default_rng(4)
    scale = 2.0 ** 40
    R, C, degree = 12, 4, 3
    X = rng.normal(0, 1, (R, C))
    wtrue = rng.uniform(-1, 1, C)
    y = (1 / (1 + np.exp(-X @ wtrue)) > rng.uniform(0, 1, R)).astype(float)
    w0 = rng.uniform(-2, 2, C)
    lay = lr.RowLayout(R, C, fx.n // 2)
    rows = np.stack([_enc(fx, 200 + i, r, scale) for i, r in enumerate(lay.rows(X))])
    cols = np.stack([_enc(fx, 300 + j, c, scale) for j, c in enumerate(lay.columns(X))])
    lab = _enc(fx, 400, lay.labels(y), scale)
    wct = _enc(fx, 401, lay.weights(w0), scale)
    br.seed = 3000
    got = lr.update_weights(fx.ev, fx.ctx.upload(rows, scale=scale), fx.ctx.upload(cols, scale=scale),
                            fx.ctx.upload(lab, scale=scale), fx.ctx.upload(wct, scale=scale), 0.1, scale,
                            fx.keys, br, br, degree=degree, method="horner")
    br.seed = 3000
    want = rw.update_weights(E, [rw.OCt(r, scale) for r in rows], [rw.OCt(c, scale) for c in cols],
                             rw.OCt(lab, scale), rw.OCt(wct, scale), 0.1, scale, lr.folded_coeffs(degree),
                             br.o_encode, br.o_encrypt, "horner")
    assert got.limbs == want.limbs and got.scale == want.scale
    assert np.array_equal(got.numpy()[0], want.data)
    dec = fx.orc.decode(fx.orc.decrypt(fx.sk, got.numpy()[0]), got.scale)[:C]
    assert np.abs(dec - lr.plain_epoch(X, y, w0, 0.1, degree)).max() < 1e-3


def test_client_objects_tolerance(make_fixture):
    """product-side CKKSEncoder / KeyGenerator / Encryptor / Decryptor (tolerance-compared):
    encode->decode, encrypt->decrypt, and a rotate + multiply/relinearize/rescale round trip with
    product-generated keys"""
    wl, lr, client = _mods()
    fx = make_fixture(13, [60, 40, 40, 60])
    ctx, ev = fx.ctx, fx.ev
    enc = client.CKKSEncoder(ctx)
    rng = np.random.default_rng(5)
    scale = 2.0 ** 40
    x, y = rng.uniform(-1, 1, (2, 200)), rng.uniform(-1, 1, (2, 200))
    px = enc.encode(x, scale)
    assert np.abs(enc.decode(px)[:, :200] - x).max() < 2.0 ** -20
    # product encoder agrees with the oracle's encoder to within rounding
    assert np.abs(fx.orc.decode(px.numpy()[0, 0], scale)[:200] - x[0]).max() < 2.0 ** -20
    kg = client.KeyGenerator(ctx, seed=7)
    keys = kg.keyset(steps=[1, -1, 4])
    encr = client.Encryptor(ctx, kg.public_key(), seed=8)
    decr = client.Decryptor(ctx, kg.secret_key())
    cx, cy = encr.encrypt(px), encr.encrypt(enc.encode(y, scale))
    assert cx.limbs == ctx.top_limbs
    assert np.abs(enc.decode(decr.decrypt(cx))[:, :200] - x).max() < 2.0 ** -20
    prod = ev.rescale_to_next(ev.relinearize(ev.multiply(cx, cy), keys))
    assert np.abs(enc.decode(decr.decrypt(prod))[:, :200] - x * y).max() < 2.0 ** -20
    rot = ev.rotate_vector(cx, 5, keys)                    # NAF: 1 + 4
    full = np.zeros((2, ctx.n // 2))
    full[:, :200] = x
    assert np.abs(enc.decode(decr.decrypt(rot)) - np.roll(full, -5, axis=1)).max() < 2.0 ** -20
    c = enc.encode(0.37, scale)
    assert np.abs(enc.decode(decr.decrypt(ev.add_plain(cx, c)))[:, :200] - (x + 0.37)).max() < 2.0 ** -20
    # encryption of a lower-level plaintext lands at that level
    low = enc.encode(x, scale, limbs=1)
    cl = encr.encrypt(low)
    assert cl.limbs == 1 and np.abs(enc.decode(decr.decrypt(cl))[:, :200] - x).max() < 2.0 ** -10


def test_train_cipher_two_iterations(eng):
    """train_cipher (logistic_regression_ckks.cpp:348-385, repair R4): two iterations with the weights
    refreshed by the key holder in between, against two steps of plaintext LR (product-side keys and
    encoder; one iteration with a degree-3 Horner sigmoid uses all 9 levels of {60, 40 x 8, 60})"""
    wl, lr, client = _mods()
    params = importlib.import_module(PKG + ".params")
    ctx = eng.Context(15, params.coeff_modulus_create(15, [60] + [40] * 8 + [60]))
    ev = eng.Evaluator(ctx)
    enc = client.CKKSEncoder(ctx)
    kg = client.KeyGenerator(ctx, seed=21)
    keys = kg.keyset(steps=[1, -4, -8])
    encr = client.Encryptor(ctx, kg.public_key(), seed=22)
    decr = client.Decryptor(ctx, kg.secret_key())
    rng = np.random.default_rng(23)
    scale, R, C, degree = 2.0 ** 40, 8, 4, 3
    X = rng.normal(0, 1, (R, C))
    y = (rng.uniform(0, 1, R) > 0.5).astype(float)
    w0 = rng.uniform(-2, 2, C)
    lay = lr.RowLayout(R, C, ctx.n // 2)
    rows = encr.encrypt(enc.encode(lay.rows(X), scale))
    cols = encr.encrypt(enc.encode(lay.columns(X), scale))
    labs = encr.encrypt(enc.encode(lay.labels(y), scale))
    wct = encr.encrypt(enc.encode(lay.weights(w0), scale))
    out = lr.train_cipher(ev, rows, cols, labs, wct, 0.1, 2, lay, scale, keys, enc, encr, decr, degree=degree)
    assert out.limbs == ctx.top_limbs                       # refreshed: back at the top level
    want = w0
    for _ in range(2):
        want = lr.plain_epoch(X, y, want, 0.1, degree)
    got = enc.decode(decr.decrypt(out))[0, :C]
    assert np.abs(got - want).max() < 1e-3


def test_column_layout_epoch_matches_plaintext(make_fixture):
    """config-5 layout at a reduced size (C = 4 features, 2 mini-batches of B = 16 samples,
    tree degree 7): decrypted gradient and updated weights vs plaintext LR"""
    wl, lr, client = _mods()
    fx = make_fixture(12, [50] + [40] * 8 + [50], steps=POW2)
    ctx, ev = fx.ctx, fx.ev
    enc = client.CKKSEncoder(ctx)
    kg = client.KeyGenerator(ctx, seed=9)
    keys = kg.keyset(steps=[1, -16])
    encr = client.Encryptor(ctx, kg.public_key(), seed=10)
    decr = client.Decryptor(ctx, kg.secret_key())
    rng = np.random.default_rng(6)
    scale = 2.0 ** 40
    C, B, M, degree = 4, 16, 2, 7
    R = B * M
    X = rng.normal(0, 1, (R, C))
    wtrue = rng.uniform(-1, 1, C)
    y = (1 / (1 + np.exp(-X @ wtrue)) > rng.uniform(0, 1, R)).astype(float)
    w0 = rng.uniform(-1, 1, C)
    lay = lr.ColumnLayout(R, C, B, ctx.n // 2)
    cols = encr.encrypt(enc.encode(lay.columns(X), scale))
    labs = encr.encrypt(enc.encode(lay.labels(y), scale))
    wb = encr.encrypt(enc.encode(np.repeat(w0[:, None], ctx.n // 2, axis=1), scale))
    wvec = np.zeros(ctx.n // 2)
    wvec[:C] = w0
    wct = encr.encrypt(enc.encode(wvec, scale))
    grad = lr.column_epoch_gradient(ev, cols, labs, wb, C, B, scale, keys, enc, encr, degree=degree, method="tree")
    g = enc.decode(decr.decrypt(grad))[0, :C]
    p = lr.sigmoid_approx(X @ w0, degree)
    assert np.abs(g - X.T @ (p - y)).max() < 1e-2
    neww = lr.apply_gradient(ev, grad, wct, 0.1, R, scale, enc)
    got = enc.decode(decr.decrypt(neww))[0, :C]
    assert np.abs(got - lr.plain_epoch(X, y, w0, 0.1, degree)).max() < 1e-3
    # SURVEY 8(f4) mode: log2(B) doubling rotations instead of B-1 unit rotations -- different
    # polynomials, same decrypted gradient within noise
    keys2 = kg.keyset(steps=[1, -16, 2, 4, 8])
    fast = lr.column_epoch_gradient(ev, cols, labs, wb, C, B, scale, keys2, enc, encr, degree=degree, method="tree",
                                    dot_method="doubling")
    assert fast.limbs == grad.limbs and fast.scale == grad.scale
    assert not np.array_equal(fast.numpy(), grad.numpy())
    assert np.abs(enc.decode(decr.decrypt(fast))[0, :C] - g).max() < 1e-5


def test_sparse_diagonal_sets_match_dense(make_fixture):
    """the de-duplicated diagonal evaluation (needed for CC_Matrix_Multiplication at d = 64) is
    bit-identical to the dense one"""
    wl, _, _ = _mods()
    fx = make_fixture(12, [50, 40, 40, 40, 40, 50], steps=POW2)
    br = Bridge(fx)
    plans = wl.PlanCache(fx.ctx, fx.keys)
    rng = np.random.default_rng(8)
    scale, eps, d = 2.0 ** 40, 1e-8, 3
    A, B = rng.uniform(0, 1, (d, d)), rng.uniform(0, 1, (d, d))
    ctA = fx.ctx.upload(_enc(fx, 500, A.reshape(-1), scale), scale=scale)
    ctB = fx.ctx.upload(_enc(fx, 501, B.reshape(-1), scale), scale=scale)

    def dense(U):
        return br.encode(wl.all_diagonals(U) + eps, scale)

    def sparse(U):
        return wl.DiagonalSet.from_matrix(U, eps, scale, br)

    mats = dict(sigma=wl.u_sigma(d), tau=wl.u_tau(d), V=[wl.v_k(d, k) for k in range(1, d)], W=[wl.w_k(d, k) for k in range(1, d)])
    want = wl.cc_matrix_multiplication(fx.ev, ctA, ctB, d, dense(mats["sigma"]), dense(mats["tau"]),
                                       [dense(m) for m in mats["V"]], [dense(m) for m in mats["W"]], fx.keys, plans)
    got = wl.cc_matrix_multiplication_sparse(fx.ev, ctA, ctB, d, sparse(mats["sigma"]), sparse(mats["tau"]),
                                             [sparse(m) for m in mats["V"]], [sparse(m) for m in mats["W"]], fx.keys, plans)
    assert got.limbs == want.limbs and got.scale == want.scale
    assert np.array_equal(got.numpy(), want.numpy())
    dec = fx.orc.decode(fx.orc.decrypt(fx.sk, got.numpy()[0]), got.scale)[: d * d].reshape(d, d)
    assert np.abs(dec - A @ B).max() < 1e-3


def test_hoisted_rotations_mode(make_fixture):
    """rotations of ONE ciphertext with a shared digit decomposition (ckks_rotate_plan_hoisted).  Not the
    reference's polynomials (SEAL permutes before lifting digits) -- checked the way north_star checks decrypted outputs:
    every hoisted rotation decrypts to the same slots as SEAL-order rotate_vector within key-switch noise (2^-19 at scale 2^40), on
    both arithmetic paths (integer limbs and FP64 limbs), at the top level and one level down, and the hoisted
    Linear_Transform_Plain / BSGS variants equal U @ v."""
    wl, _, client = _mods()
    steps = [-16] + list(range(1, 16)) + [16, 32, 48]
    fx = make_fixture(13, [60, 40, 40, 60], steps=tuple(steps))
    plans = wl.PlanCache(fx.ctx, fx.keys)
    enc = client.CKKSEncoder(fx.ctx)
    rng = np.random.default_rng(40)
    scale = 2.0 ** 40
    x = rng.uniform(-1, 1, 64)
    full = np.zeros(fx.n // 2)
    full[:64] = x
    for L in (fx.L, fx.L - 1):
        ct = fx.ctx.upload(_enc(fx, 700 + L, x, scale)[:, :L], cap=fx.L, scale=scale)
        plan = plans.get([0, 1, 2, 3, 5, 7, 11, 15, -16])
        ref = fx.ev.rotate_plan(ct, plan)
        got = fx.ev.rotate_plan_hoisted(ct, plan)
        assert got.limbs == L and got.batch == plan.batch
        assert np.array_equal(got.numpy()[0], ct.numpy()[0])                     # rotate by 0 = the input
        assert not np.array_equal(got.numpy()[1:], ref.numpy()[1:])              # different polynomials ...
        for b, st in enumerate(plan.steps):                                     # ... same plaintext
            dg = fx.orc.decode(fx.orc.decrypt(fx.sk, got.numpy()[b]), scale)
            dr = fx.orc.decode(fx.orc.decrypt(fx.sk, ref.numpy()[b]), scale)
            # key-switch noise at N = 8192 with one 60-bit special prime is ~1e-6 on slots of magnitude 1 (the SEAL-order
            # rotation shows the same): bound 2^-19 against the exact rotation, 2^-18 between the two noisy results
            assert np.abs(dg - np.roll(full, -st)).max() < 2.0 ** -19, (L, st)
            assert np.abs(dr - np.roll(full, -st)).max() < 2.0 ** -19, (L, st)
            assert np.abs(dg - dr).max() < 2.0 ** -18, (L, st)
    # a step without its own key cannot be hoisted (it would be a NAF chain)
    capi = importlib.import_module(PKG + ".capi")
    ct = fx.ctx.upload(_enc(fx, 710, x, scale), scale=scale)
    with pytest.raises(capi.CkksInvalidArgument, match="key"):
        fx.ev.rotate_plan_hoisted(ct, plans.get([1, 17]))
    # Linear_Transform_Plain, hoisted and BSGS + hoisted baby steps
    d = 16
    U, v = rng.uniform(0, 1, (d, d)), rng.uniform(0, 1, d)
    ctv = fx.ctx.upload(_enc(fx, 720, v, scale), scale=scale)
    diags = enc.encode(wl.all_diagonals(U), scale)
    ref = wl.linear_transform_plain(fx.ev, ctv, diags, fx.keys, plans)
    got = wl.linear_transform_plain_hoisted(fx.ev, ctv, diags, fx.keys, plans)
    assert got.limbs == ref.limbs and got.scale == ref.scale and not np.array_equal(got.numpy(), ref.numpy())
    dg = fx.orc.decode(fx.orc.decrypt(fx.sk, got.numpy()[0]), got.scale)[:d]
    dr = fx.orc.decode(fx.orc.decrypt(fx.sk, ref.numpy()[0]), ref.scale)[:d]
    assert np.abs(dg - dr).max() < 1e-5 and np.abs(dg - U @ v).max() < 1e-4
    d = 64
    U, v = rng.uniform(0, 1, (d, d)), rng.uniform(0, 1, d)
    fx2 = make_fixture(13, [60, 40, 40, 60], steps=tuple([-64] + list(range(1, 16)) + [16, 32, 48]))
    plans2 = wl.PlanCache(fx2.ctx, fx2.keys)
    enc2 = client.CKKSEncoder(fx2.ctx)
    ctv = fx2.ctx.upload(_enc(fx2, 721, v, scale), scale=scale)
    bd = wl.BsgsDiagonals(U, scale, enc2, baby=16)
    got = wl.linear_transform_plain_bsgs_hoisted(fx2.ev, ctv, bd, fx2.keys, plans2)
    dg = fx2.orc.decode(fx2.orc.decrypt(fx2.sk, got.numpy()[0]), got.scale)[:d]
    assert np.abs(dg - U @ v).max() < 1e-4
