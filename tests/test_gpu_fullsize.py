"""GPU parity at the sizes BASELINE.json names, bit-exact against
the sequential oracle where the oracle finishes in seconds."""
import importlib

import numpy as np
import pytest

import ref_workloads as rw

pytestmark = pytest.mark.gpu
PKG = "seal-fyp-logistic-regression_b200"
POW2_14 = tuple(s for i in range(13) for s in (1 << i, -(1 << i)))


@pytest.mark.parametrize("d", [64, 128])
def test_config3_linear_transform_n16384(make_fixture, d):
    """config 3: Linear_Transform_Plain, N = 16384, {60,40,40,60}, scale 2^40, d = 64 / 128, entries
    uniform [0,1) from a fixed seed; 157 / 356 key switches with SEAL's default Galois keys"""
    wl = importlib.import_module(PKG + ".workloads")
    fx = make_fixture(14, [60, 40, 40, 60], steps=POW2_14)
    E = rw.OEval(fx.orc, fx.rlk, fx.gks)
    plans = wl.PlanCache(fx.ctx, fx.keys)
    rng = np.random.default_rng(d)
    scale = 2.0 ** 40
    U, v = rng.uniform(0, 1, (d, d)), rng.uniform(0, 1, d)
    diags = wl.all_diagonals(U)
    ct = fx.orc.encrypt(5, fx.pk, fx.orc.encode(v, scale))
    pts = np.stack([fx.orc.encode(dg, scale) for dg in diags])
    want = rw.linear_transform_plain(E, rw.OCt(ct, scale), [rw.OCt(p, scale) for p in pts])
    got = wl.linear_transform_plain(fx.ev, fx.ctx.upload(ct, scale=scale), fx.ctx.upload_plain(pts, scale=scale), fx.keys, plans)
    assert plans.get(range(d)).keyswitches + len(__import__("oracle.pyoracle", fromlist=["naf"]).naf(-d)) == {64: 157, 128: 356}[d]
    assert plans.get(range(d)).keyswitches_shared == {64: 84, 128: 169}[d]    # what actually runs: common NAF prefixes once
    assert np.array_equal(got.numpy()[0], want.data)
    dec = fx.orc.decode(fx.orc.decrypt(fx.sk, got.numpy()[0]), got.scale)[:d]
    assert np.abs(dec - U @ v).max() < 1e-3


@pytest.mark.parametrize("log_n", [12, 13, 14])
def test_config2_op_sweep_bfv_default_chains(po, eng, log_n):
    """config 2: the benchmark.cpp parameter sets (CoeffModulus::BFVDefault, scale = sqrt(last prime),
    inputs v1[i] = i, v2[i] = (i % 2) + 1, benchmark.cpp:137,225-273): multiply_plain, multiply +
    relinearize, rotate by 1 and by a 3-term NAF step, rescale -- at the top data level"""
    primes = po.bfv_default(log_n)
    orc = po.Oracle(log_n, primes)
    ctx = eng.Context(log_n, primes)
    ev = eng.Evaluator(ctx)
    sk = orc.gen_secret(1)
    pk = orc.gen_public(2, sk)
    rlk = orc.gen_relin_key(3, sk)
    gks = orc.gen_galois_keys(4, sk, steps=[1, -1, -4, 16])
    keys = eng.KeySet(ctx)
    keys.set_relin(ctx.upload_key(rlk))
    for g, k in gks.items():
        keys.set_galois(g, ctx.upload_key(k))
    scale = float(np.sqrt(primes[-1]))
    for size in (10, 100, 1000):
        v1 = np.arange(size, dtype=float)
        v2 = (np.arange(size) % 2 + 1).astype(float)
        p1, p2 = orc.encode(v1, scale), orc.encode(v2, scale)
        c1, c2 = orc.encrypt(10 + size, pk, p1), orc.encrypt(11 + size, pk, p2)
        d1, d2 = ctx.upload(c1, scale=scale), ctx.upload(c2, scale=scale)
        dp2 = ctx.upload_plain(p2, scale=scale)
        assert np.array_equal(ev.add_plain(d1, dp2).numpy()[0], orc.add_plain(c1, p2))
        assert np.array_equal(ev.add(d1, d2).numpy()[0], orc.add(c1, c2))
        assert np.array_equal(ev.multiply_plain(d1, dp2).numpy()[0], orc.multiply_plain(c1, p2))
        m = ev.multiply(d1, d2)
        om = orc.multiply(c1, c2)
        assert np.array_equal(m.numpy()[0], om)
        r = ev.relinearize(m, keys)
        orl = orc.relinearize(om, rlk)
        assert np.array_equal(r.numpy()[0], orl)
        assert np.array_equal(ev.rescale_to_next(r).numpy()[0], orc.rescale(orl))
        assert np.array_equal(ev.rotate_vector(d1, 1, keys).numpy()[0], orc.rotate(c1, 1, gks))
        assert po.naf(11) == [-1, -4, 16]
        assert np.array_equal(ev.rotate_vector(d1, 11, keys).numpy()[0], orc.rotate(c1, 11, gks))
        dec = orc.decode(orc.decrypt(sk, r.numpy()[0]), r.scale)[:size]
        assert np.abs(dec - v1 * v2).max() < 1e-2 * max(1.0, size)
