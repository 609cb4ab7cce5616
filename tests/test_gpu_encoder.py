"""GPU parity, SURVEY.md section 8 row f1: CKKSEncoder::encode / decode on the device
(include/ckks_b200.h: ckks_encode, ckks_encode_scalar, ckks_decode) against the CPU oracle's encoder.

north_star compares encode/decode by tolerance (|err| < 2^-20 relative at scale 2^40), not bit-exactly:
both sides run a floating-point DFT.  What IS checked exactly: coefficients after rounding differ by
at most one unit, the batched call equals single calls bit for bit, limbs of one plaintext are
residues of the same integer, and scalar encodes equal the oracle's."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PKG = "seal-fyp-logistic-regression_b200"
TOL = 2.0 ** -20          # the tolerance north_star states, at scale 2^40


def _torch():
    import torch
    return torch


def _centered(res, p):
    r = res.astype(np.int64)
    return np.where(r > p // 2, r - p, r)


@pytest.mark.parametrize("log_n,bits", [(12, [36, 36, 37]), (13, [60, 40, 40, 60]), (14, [60, 40, 40, 40, 40, 60]),
                                        (15, [60, 40, 40, 60])])
def test_encode_matches_oracle_coefficients(make_fixture, log_n, bits):
    """device encode vs oracle encode: the rounded coefficients agree to +-1 and every limb holds the
    same integer; oracle decode of the device plaintext returns the input"""
    torch = _torch()
    fx = make_fixture(log_n, bits, steps=(1,))
    n, scale = fx.ctx.n, 2.0 ** 30 if log_n == 12 else 2.0 ** 40
    rng = np.random.default_rng(log_n)
    for count in (1, 7, n // 4, n // 2):
        x = rng.uniform(-4, 4, count)
        got = fx.ev.encode(torch.from_numpy(x), scale)          # top data level
        L = got.limbs
        want = fx.orc.encode(x, scale, L)
        g = got.numpy()[0, 0]
        cg = np.stack([_centered(fx.orc.intt(j, g[j]), fx.primes[j]) for j in range(L)])
        cw = np.stack([_centered(fx.orc.intt(j, want[j]), fx.primes[j]) for j in range(L)])
        assert np.abs(cg[0] - cw[0]).max() <= 1
        assert all(np.array_equal(cg[0], cg[j]) for j in range(1, L))     # same integer in every limb
        assert np.abs(fx.orc.decode(g, scale)[:count] - x).max() < TOL
        if count < n // 2:
            assert np.abs(fx.orc.decode(g, scale)[count:]).max() < TOL    # unused slots are zero


@pytest.mark.parametrize("log_n,bits", [(12, [36, 36, 37]), (14, [60, 40, 40, 40, 40, 60]), (15, [60, 40, 40, 60])])
def test_decode_matches_oracle(make_fixture, log_n, bits):
    """device decode vs oracle decode at every level, on plaintexts with negative and large coefficients
    (a decrypted ciphertext: message + noise), batched"""
    fx = make_fixture(log_n, bits, steps=(1,))
    n, scale = fx.ctx.n, 2.0 ** 30 if log_n == 12 else 2.0 ** 40
    rng = np.random.default_rng(log_n + 50)
    x = rng.uniform(-10, 10, (3, n // 2))
    for L in range(1, fx.L + 1):
        pts = np.stack([fx.orc.decrypt(fx.sk, fx.orc.encrypt(9 + b, fx.pk, fx.orc.encode(x[b], scale)))[:L] for b in range(3)])
        dev = fx.ctx.upload_plain(pts, cap=fx.L, scale=scale)     # capacity above the level: strided view
        got = fx.ev.decode(dev).cpu().numpy()
        want = np.stack([fx.orc.decode(pts[b], scale) for b in range(3)])
        assert got.shape == want.shape
        assert np.abs(got - want).max() < 1e-9 * max(1.0, np.abs(want).max())
        if L > 1 or log_n != 12:
            assert np.abs(got - x).max() < 1e-4                   # and it is the message (+ noise / scale)
        assert np.array_equal(dev.numpy()[:, 0], pts)             # the source plaintext is untouched


def test_batch_equals_single_and_roundtrip(make_fixture):
    torch = _torch()
    fx = make_fixture(13, [60, 40, 40, 60], steps=(1,))
    rng = np.random.default_rng(3)
    scale = 2.0 ** 40
    x = rng.normal(0, 1, (5, 300))
    batch = fx.ev.encode(torch.from_numpy(x), scale, limbs=2, cap=3)
    for b in range(5):
        one = fx.ev.encode(torch.from_numpy(x[b]), scale, limbs=2)
        assert np.array_equal(batch.numpy()[b], one.numpy()[0])
    back = fx.ev.decode(batch).cpu().numpy()
    assert np.abs(back[:, :300] - x).max() < TOL and np.abs(back[:, 300:]).max() < TOL
    # small workspace cap: the batch is processed in chunks, results unchanged
    fx.ctx.set_workspace_cap(2 * fx.ctx.n * 16)
    again = fx.ev.encode(torch.from_numpy(x), scale, limbs=2, cap=3)
    assert np.array_equal(again.numpy(), batch.numpy())
    assert np.array_equal(fx.ev.decode(again).cpu().numpy(), back)
    fx.ctx.set_workspace_cap(1 << 30)


def test_scalar_encode_and_large_scales(make_fixture):
    """encode(double) equals the oracle's; scales beyond 64 bits take the mantissa/exponent path"""
    torch = _torch()
    fx = make_fixture(13, [60, 40, 40, 60], steps=(1,))
    for value, scale in ((0.37, 2.0 ** 40), (-1.20069, 2.0 ** 40), (0.5, 2.0 ** 80), (-3.25, 2.0 ** 100), (0.0, 2.0 ** 40)):
        got = fx.ev.encode_scalar(value, scale, limbs=3, batch=2).numpy()
        want = fx.orc.encode(value, scale, 3)
        assert np.array_equal(got[0, 0], want) and np.array_equal(got[1, 0], want)
    rng = np.random.default_rng(11)
    x = rng.uniform(-1, 1, 1000)
    for scale in (2.0 ** 80, 2.0 ** 120):                  # coefficients far beyond 2^63
        pt = fx.ev.encode(torch.from_numpy(x), scale, limbs=3)
        assert np.abs(fx.orc.decode(pt.numpy()[0, 0], scale)[:1000] - x).max() < 2.0 ** -30
        assert np.abs(fx.ev.decode(pt).cpu().numpy()[0, :1000] - x).max() < 2.0 ** -30


def test_encoder_errors(make_fixture):
    torch = _torch()
    pkg = importlib.import_module(PKG)
    capi = importlib.import_module(PKG + ".capi")
    fx = make_fixture(12, [36, 36, 37], steps=(1,))
    n = fx.ctx.n
    with pytest.raises(capi.CkksInvalidArgument):
        fx.ev.encode(torch.zeros(n // 2 + 1, dtype=torch.float64), 2.0 ** 30)
    with pytest.raises(capi.CkksInvalidArgument):
        fx.ev.encode(torch.zeros(4, dtype=torch.float64), 0.0)
    with pytest.raises(capi.CkksInvalidArgument):
        fx.ev.decode(fx.ctx.empty(1, 2, 2))                      # a ciphertext is not a plaintext
    # an empty value vector encodes the zero plaintext (SEAL accepts it)
    z = fx.ev.encode(torch.zeros((1, 0), dtype=torch.float64), 2.0 ** 30)
    assert not z.numpy().any()


def test_device_sampler_distributions(make_fixture):
    """ckks_sample: ternary / clipped normal / uniform polynomials drawn on the device.
    Statistical checks (the generator is not SEAL's, outputs are compared by distribution), exact checks
    for structure: the same small integer in every limb, range, determinism per (seed, stream)."""
    fx = make_fixture(13, [60, 40, 40, 60], steps=(1,))
    ev, n, K = fx.ev, fx.ctx.n, len(fx.primes)

    def small(kind, seed, stream, count=4):
        t = ev.sample(kind, seed, stream, count, K).cpu().numpy().view(np.uint64)
        cent = np.stack([[_centered(fx.orc.intt(j, t[b, j]), fx.primes[j]) for j in range(K)] for b in range(count)])
        assert all(np.array_equal(cent[:, 0], cent[:, j]) for j in range(1, K))      # one integer, K residues
        return cent[:, 0]

    tern = small(ev.TERNARY, 1234, 1)
    assert set(np.unique(tern)) == {-1, 0, 1}
    for v in (-1, 0, 1):
        assert abs((tern == v).mean() - 1 / 3) < 0.01
    err = small(ev.NORMAL, 1234, 2)
    assert np.abs(err).max() <= 19
    assert abs(err.mean()) < 0.1 and abs(err.std() - np.sqrt(3.2 ** 2 + 1 / 12)) < 0.1
    assert np.array_equal(small(ev.NORMAL, 1234, 2), err)                            # reproducible
    assert not np.array_equal(small(ev.NORMAL, 1234, 3), err)                        # new stream, new draw
    assert not np.array_equal(small(ev.NORMAL, 1235, 2), err)                        # new seed, new draw
    assert not np.array_equal(err[0], err[1])                                        # polynomials differ
    uni = ev.sample(ev.UNIFORM, 99, 7, 8, K).cpu().numpy().view(np.uint64)
    for j, p in enumerate(fx.primes):
        assert uni[:, j].max() < p
        assert abs(uni[:, j].astype(np.float64).mean() / p - 0.5) < 0.01
        assert abs(uni[:, j].astype(np.float64).std() / p - np.sqrt(1 / 12)) < 0.01
    assert not np.array_equal(uni[:, 1] % 1000003, uni[:, 2] % 1000003)              # limbs are independent
