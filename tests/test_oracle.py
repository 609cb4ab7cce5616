"""CPU tests: pin the oracle against everything that can be pinned without SEAL itself
(the reference holds no golden vectors; parity is otherwise unpinned)."""
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_bfv_default_tables_are_ntt_primes(po):
    # SEAL's hard-coded 128-bit default moduli (benchmark.cpp:137): prime, 1 mod 2N, bit totals
    # equal CoeffModulus::MaxBitCount (README.md:176 quotes 218 for N = 8192)
    for log_n, count, total in ((12, 3, 109), (13, 5, 218), (14, 9, 438), (15, 16, 881)):
        ps = po.bfv_default(log_n)
        assert len(ps) == count
        assert sum(p.bit_length() for p in ps) == total == po.max_bit_count(log_n)
        assert all(po.is_prime(p) and (p - 1) % (2 << log_n) == 0 for p in ps)
        assert len(set(ps)) == count


def test_coeff_modulus_create_golden(po):
    gold = json.load(open(os.path.join(GOLD, "coeff_modulus.json")))
    for case in gold["create"]:
        got = po.coeff_modulus_create(case["log_n"], case["bits"])
        assert [hex(p) for p in got] == case["primes"], case
        for p, b in zip(got, case["bits"]):
            assert p.bit_length() == b and po.is_prime(p) and (p - 1) % (2 << case["log_n"]) == 0
        # first occurrence of a size gets the smallest prime of that size, the last the largest
        by_size = {}
        for p, b in zip(got, case["bits"]):
            by_size.setdefault(b, []).append(p)
        assert all(v == sorted(v) for v in by_size.values())


def test_naf_golden(po):
    # examples worked out in SURVEY.md A.6 from SEAL's util::naf
    assert po.naf(3) == [-1, 4]
    assert po.naf(7) == [-1, 8]
    assert po.naf(100) == [4, -32, 128]
    assert po.naf(1) == [1] and po.naf(-64) == [-64]
    for v in list(range(-300, 300)) + [4095, 8191, -8191]:
        parts = po.naf(v)
        assert sum(parts) == v
        mags = [abs(x) for x in parts]
        assert all(m & (m - 1) == 0 for m in mags) and mags == sorted(mags)
        # non-adjacent: no two consecutive powers of two
        assert all(b >= 4 * a for a, b in zip(mags, mags[1:]))


def test_keyswitch_counts_match_survey(po):
    # key switches of Linear_Transform_Plain at dimension d with default keys
    def ks(d):
        return len(po.naf(-d)) + sum(len(po.naf(l)) for l in range(1, d))
    assert ks(10) == 16 and ks(64) == 157 and ks(100) == 270 and ks(128) == 356


def test_ntt_matches_definition(po):
    o = po.Oracle(5, po.coeff_modulus_create(5, [30, 31]))
    rng = np.random.default_rng(1)
    for j in range(2):
        a = rng.integers(0, o.primes[j], size=32, dtype=np.uint64)
        assert np.array_equal(o.ntt(j, a), o.ntt_naive(j, a))
        assert np.array_equal(o.intt(j, o.ntt(j, a)), a)
        # minimal primitive root: primitive 2N-th root, and no smaller one exists
        psi, p = o.psi(j), o.primes[j]
        assert pow(psi, 32, p) == p - 1
        assert all(pow(x, 32, p) != p - 1 for x in range(2, min(psi, 5000)))


def test_ntt_negacyclic_convolution(po):
    log_n = 12
    o = po.Oracle(log_n, po.coeff_modulus_create(log_n, [40, 40]))
    rng = np.random.default_rng(2)
    n, p = o.n, o.primes[0]
    a = np.zeros(n, dtype=np.uint64)
    b = np.zeros(n, dtype=np.uint64)
    ia, ib = rng.integers(0, n, 5), rng.integers(0, n, 5)
    a[ia] = rng.integers(1, p, 5, dtype=np.uint64)
    b[ib] = rng.integers(1, p, 5, dtype=np.uint64)
    want = [0] * n
    for i in np.nonzero(a)[0]:
        for k in np.nonzero(b)[0]:
            v = int(a[i]) * int(b[k]) % p
            d = int(i + k)
            if d >= n:
                want[d - n] = (want[d - n] - v) % p
            else:
                want[d] = (want[d] + v) % p
    A = o.ntt(0, a)[None, None]
    B = o.ntt(0, b)[None]
    prod = np.array([(int(x) * int(y)) % p for x, y in zip(A[0, 0], B[0])], dtype=np.uint64)
    assert np.array_equal(o.intt(0, prod), np.array(want, dtype=np.uint64))


@pytest.fixture(scope="module")
def small(po):
    log_n = 12
    primes = po.coeff_modulus_create(log_n, [50, 40, 40, 50])
    o = po.Oracle(log_n, primes)
    sk = o.gen_secret(1)
    return dict(o=o, sk=sk, pk=o.gen_public(2, sk), rlk=o.gen_relin_key(3, sk), gks=o.gen_galois_keys(4, sk),
                scale=2.0 ** 40, primes=primes)


def _dec(s, ct, scale, n=64):
    return s["o"].decode(s["o"].decrypt(s["sk"], ct), scale)[:n]


def test_oracle_homomorphic_semantics(po, small):
    """decrypt(eval(enc(x))) == plaintext math: the plaintext checks the reference prints
    (linear_transformation.cpp:203-218, polynomial.cpp:171-204) applied to every evaluator op"""
    o, s = small["o"], small
    rng = np.random.default_rng(3)
    x, y = rng.uniform(-1, 1, 64), rng.uniform(-1, 1, 64)
    sc = s["scale"]
    cx, cy = o.encrypt(10, s["pk"], o.encode(x, sc)), o.encrypt(11, s["pk"], o.encode(y, sc))
    assert np.abs(_dec(s, cx, sc) - x).max() < 1e-7
    assert np.abs(_dec(s, o.add(cx, cy), sc) - (x + y)).max() < 1e-7
    assert np.abs(_dec(s, o.sub(cx, cy), sc) - (x - y)).max() < 1e-7
    assert np.abs(_dec(s, o.negate(cx), sc) + x).max() < 1e-7
    m3 = o.multiply(cx, cy)
    assert m3.shape[0] == 3 and np.abs(_dec(s, m3, sc * sc) - x * y).max() < 1e-7
    m2 = o.relinearize(m3, s["rlk"])
    assert np.abs(_dec(s, m2, sc * sc) - x * y).max() < 1e-6
    r = o.rescale(m2)
    assert r.shape[1] == 2 and np.abs(_dec(s, r, sc * sc / s["primes"][2]) - x * y).max() < 1e-6
    mp = o.multiply_plain(cx, o.encode(y, sc))
    assert np.abs(_dec(s, mp, sc * sc) - x * y).max() < 1e-7
    ap = o.add_plain(cx, o.encode(0.37, sc))
    assert np.abs(_dec(s, ap, sc) - (x + 0.37)).max() < 1e-7
    ms = o.mod_switch(cx)
    assert np.abs(_dec(s, ms, sc) - x).max() < 1e-7
    # symmetric encryption decrypts too
    assert np.abs(_dec(s, o.encrypt_symmetric(5, s["sk"], o.encode(x, sc)), sc) - x).max() < 1e-7


def test_oracle_rotation_and_naf_chain(po, small):
    o, s = small["o"], small
    rng = np.random.default_rng(4)
    x = rng.uniform(-1, 1, 100)
    full = np.zeros(o.n // 2)
    full[:100] = x
    sc = s["scale"]
    cx = o.encrypt(12, s["pk"], o.encode(x, sc))
    for st in (1, -1, 3, 7, 100, -100, 2047):
        got = o.decode(o.decrypt(s["sk"], o.rotate(cx, st, s["gks"])), sc)
        assert np.abs(got - np.roll(full, -st)).max() < 1e-5, st
    with pytest.raises(ValueError):
        o.galois_elt(o.n // 2)
    only_one = {o.galois_elt(1): s["gks"][o.galois_elt(1)]}
    with pytest.raises(KeyError):
        o.rotate(cx, 4, only_one)


def test_rescale_rounding_switch(po, small):
    """rounding vs flooring differ by at most one unit in the last place"""
    o, s = small["o"], small
    rng = np.random.default_rng(5)
    ct = np.stack([rng.integers(0, p, size=(2, o.n), dtype=np.uint64) for p in s["primes"][:3]], axis=1)
    a = o.rescale(ct)
    o.set_rounding(False)
    b = o.rescale(ct)
    o.set_rounding(True)
    assert not np.array_equal(a, b)
    # compare after INTT: coefficients differ by 0 or 1 (mod q)
    for j in range(2):
        da = o.intt(j, a[0, j]).astype(object)
        db = o.intt(j, b[0, j]).astype(object)
        diff = (da - db) % s["primes"][j]
        assert set(np.unique(diff).tolist()) <= {0, 1}


def test_transparent_detection(po, small):
    o, s = small["o"], small
    ct = np.zeros((2, 3, o.n), dtype=np.uint64)
    ct[0] = 5
    assert o.is_transparent(ct)
    ct[1, 2, 17] = 1
    assert not o.is_transparent(ct)


def test_golden_ciphertext_digest(po):
    """regression pin of the oracle's own outputs (generated by tests/golden/make_golden.py):
    guards the restatement against accidental change; it is NOT a SEAL vector."""
    import hashlib
    gold = json.load(open(os.path.join(GOLD, "oracle_digests.json")))
    from golden.make_golden import compute_digests
    assert compute_digests(po) == gold
